"""CPU, only where the compiled reference exists (oracle/_ref, built from /root/reference by oracle/build_ref.py): the C
oracle against the UNMODIFIED reference, stage by stage and over the VPP flag space.  Skipped on boxes without it; the
committed golden vectors (test_oracle_golden.py) carry the same pin there."""
import itertools

import numpy as np
import pytest

from conftest import assert_same

ref = pytest.importorskip("oracle.ref")
if not ref.available():
    pytest.skip("oracle/_ref not built on this box", allow_module_level=True)


@pytest.fixture(scope="module")
def R():
    return ref.load_pinned()


@pytest.mark.parametrize("shape,D,channels", [((64, 96), 32, 1), ((50, 121), 64, 3), ((40, 70), 72, 3), ((33, 200), 192, 1)])
def test_rsgm_stages_vs_reference(orc, R, shape, D, channels):
    from vppstereo_b200 import synth
    m = R.rsgm
    p = synth.make_pair(D, shape=shape, hints="random", channels=channels)
    H0, W0 = shape
    Hp, Wp = (H0 + 15) // 16 * 16, (W0 + 15) // 16 * 16
    pad = lambda a: orc.pad_reflect(a, (Hp - H0) // 2, Hp - H0 - (Hp - H0) // 2, (Wp - W0) // 2, Wp - W0 - (Wp - W0) // 2)
    L, Rt = pad(p["left"]), pad(p["right"])
    ctl, ctr = m._census_transform(L, Rt)
    gl, gr = (orc.rgb2gray(a) if a.ndim == 3 else a for a in (L, Rt))
    ocl = np.zeros_like(ctl); ocr = np.zeros_like(ctr)
    orc.census5x5_SSE(gl, ocl, Wp, Hp); orc.census5x5_SSE(gr, ocr, Wp, Hp)
    assert_same(ocl, ctl, "census L"); assert_same(ocr, ctr, "census R")
    dsi = m._hamming_matching(ctl, ctr, D)
    odsi = np.zeros_like(dsi); orc.costMeasureCensus5x5_xyd_SSE(ocl, ocr, odsi, Wp, Hp, D, 1)
    assert_same(odsi, dsi, "cost volume")
    agg = m._aggregate_dsi(L, dsi, 11, 17, 0.5, 35)
    oagg = np.zeros_like(agg); orc.aggregate_SSE(L, odsi, oagg, Wp, Hp, D, 11, 17, 0.5, 35)
    assert_same(oagg, agg, "aggregated volume (arguments ignored)")
    for f_ref, f_orc, nm in ((R.pyrSGM.matchWTA_SSE, orc.matchWTA_SSE, "WTA L"), (R.pyrSGM.matchWTARight_SSE, orc.matchWTARight_SSE, "WTA R")):
        a = np.zeros((Hp, Wp), np.float32); f_ref(agg, a, Wp, Hp, D, 0.95)
        b = np.zeros((Hp, Wp), np.float32); f_orc(agg, b, Wp, Hp, D, 0.95)
        assert_same(b, a, nm)
    for method in (0, 1):
        a = np.zeros((Hp, Wp), np.float32); R.pyrSGM.matchWTA_SSE(agg, a, Wp, Hp, D, 0.95); b = a.copy()
        R.pyrSGM.subPixelRefine(agg, a, Wp, Hp, D, method); orc.subPixelRefine(agg, b, Wp, Hp, D, method)
        assert_same(b, a, f"subPixelRefine({method})")
    med = np.zeros((Hp, Wp), np.float32); R.pyrSGM.median3x3_SSE(a, med, Wp, Hp)
    omed = np.zeros((Hp, Wp), np.float32); orc.median3x3_SSE(a, omed, Wp, Hp)
    assert_same(omed, med, "median")
    assert_same(orc.compute_rsgm(p["left"], p["left"], p["right"], dmax=D), m.compute_rsgm(p["left"], p["left"], p["right"], dmax=D),
                "compute_rsgm")
    h = p["hints"]; v = (h > 0).astype(np.float32)
    assert_same(orc.compute_rsgm(p["left"], p["left"], p["right"], hints=h, validhints=v, dmax=D, subpixel=False),
                m.compute_rsgm(p["left"], p["left"], p["right"], hints=h, validhints=v, dmax=D, subpixel=False), "guided compute_rsgm")


def test_vpp_flag_space_vs_reference(orc, R):
    from numba import njit
    from vppstereo_b200 import synth

    @njit
    def nb_seed(s):
        np.random.seed(s)

    @njit
    def nb_draw(n):
        out = np.empty(n, np.uint8)
        for i in range(n):
            out[i] = np.random.randint(0, 256)
        return out

    V, S = R.vpp_core_opt, R.vpp_standalone
    rng = np.random.default_rng(5)
    combos = [c for i, c in enumerate(itertools.product([1, 3], [1, 3, 5], [0, 1], [0, 1], [0, 1], [0, 1], [0, 1], [0, 1])) if i % 9 == 0]
    H, W = 36, 64
    for (C, wsize, direction, uniform, interp, discard, occ_on, dark) in combos:
        p = synth.make_pair(int(rng.integers(100)), shape=(H, W), hints="random", channels=3, density=0.08)
        l0 = p["left"][..., :C].copy(); r0 = p["right"][..., :C].copy()
        if dark:
            l0 //= 100; r0 //= 100
        g = (p["hints"] * 0.2).astype(np.float32)
        g[5, 3] = 7.0; g[6, 60] = 30.5; g[0, 0] = 2.5; g[H - 1, W - 1] = 1.0
        g_occ = ((rng.random((H, W)) < 0.3) & (occ_on == 1)).astype(np.uint8)
        aggx, aggy = (64, 5) if dark else (16, 3)
        n = orc.stream_length(g, wsize, C, uniform)
        st = orc.libc_rand_stream(7, n)
        la, ra = l0.copy(), r0.copy(); V.init_rand(7)
        na = V.virtual_projection_scan_rnd(la, ra, g, W, H, C, uniform, wsize, direction, 0.4, 0.15, g_occ, discard, interp)
        lb, rb = l0.copy(), r0.copy()
        nb = orc.virtual_projection_scan_rnd(lb, rb, g, W, H, C, uniform, wsize, direction, 0.4, 0.15, g_occ, discard, interp, stream=st, mode=0)
        assert na == nb
        assert_same(lb, la, "cython rnd L"); assert_same(rb, ra, "cython rnd R")
        la, ra = l0.copy(), r0.copy()
        V.virtual_projection_scan_max_dist(la, ra, g, W, H, C, uniform, wsize, aggx, aggy, direction, 0.4, 0.15, g_occ, discard, interp)
        lb, rb = l0.copy(), r0.copy()
        orc.virtual_projection_scan_max_dist(lb, rb, g, W, H, C, uniform, wsize, aggx, aggy, direction, 0.4, 0.15, g_occ, discard, interp, mode=0)
        assert_same(lb, la, "cython maxDistance L"); assert_same(rb, ra, "cython maxDistance R")
        nb_seed(11); stn = nb_draw(n); nb_seed(11)
        kw = dict(wsize=wsize, wsizeAgg_x=aggx, wsizeAgg_y=aggy, left2right=bool(direction), blending=0.4, uniform_color=bool(uniform),
                  c_occ=0.15, discard_occ=bool(discard), interpolate=bool(interp))
        li = l0 if C == 3 else l0[..., 0]; ri = r0 if C == 3 else r0[..., 0]
        la, ra = S.vpp(li, ri, g, method="rnd", g_occ=g_occ.astype(np.float32), **kw)
        lb, rb = orc.vpp(li, ri, g, method="rnd", stream=stn, mode=1, g_occ=g_occ, **kw)
        assert_same(lb, la, "numba rnd L"); assert_same(rb, ra, "numba rnd R")
        la, ra = S.vpp(li, ri, g, method="maxDistance", g_occ=g_occ.astype(np.float32), **kw)
        lb, rb = orc.vpp(li, ri, g, method="maxDistance", mode=1, g_occ=g_occ, **kw)
        assert_same(lb, la, "numba maxDistance L"); assert_same(rb, ra, "numba maxDistance R")


@pytest.mark.parametrize("shape,kind,density,fg,kw", [
    ((96, 200), "lidar", 0.05, 0, {}), ((120, 160), "random", 0.05, 4, {}), ((64, 64), "random", 0.4, 2, {}),
    ((80, 140), "random", 0.1, 3, dict(rx=5, ry=11, l=1, g=0.3, th_conf=2)), ((375, 1242), "lidar", 0.05, 6, {}),
    ((50, 90), "random", 0.2, 3, dict(th_filter=2)),
])
def test_occlusion_heuristic_vs_reference(orc, R, shape, kind, density, fg, kw):
    """filter.py:246-292 (numba) against oracle/filter_oracle.c: mask and filtered hints, bit for bit."""
    from vppstereo_b200 import synth
    g = synth.make_pair(7, shape=shape, hints=kind, density=density, foreground=fg)["hints"].astype(np.float32)
    d_ref, c_ref = R.filter.occlusion_heuristic(g.copy(), **kw)
    d, c = orc.occlusion_heuristic(g, **kw)
    assert c_ref.dtype == np.uint8 and d_ref.dtype == np.float32
    assert_same(c, c_ref, "occlusion mask"); assert_same(d, d_ref, "filtered hints")


@pytest.mark.parametrize("C,wsize,distance,bilateral,uniform,method", [
    (3, 7, 1, 1, 0, "rnd"), (3, 5, 0, 1, 0, "rnd"), (1, 7, 1, 0, 1, "rnd"), (3, 9, 1, 1, 1, "rnd"),
    (3, 7, 1, 1, 0, "maxDistance"), (1, 5, 0, 1, 1, "maxDistance"), (3, 5, 1, 0, 0, "maxDistance"),
])
def test_vpp_adaptive_patches_vs_reference(orc, R, C, wsize, distance, bilateral, uniform, method):
    """TPAMI extensions of vpp() (distance-based patch size vpp_standalone.py:6-11, bilateral adaptive patch :371-394 and
    the guards :153-154/:334-335): oracle against the reference's numba code, bit for bit, with numba's own random stream."""
    from numba import njit
    from vppstereo_b200 import synth

    @njit
    def nb_seed(s):
        np.random.seed(s)

    @njit
    def nb_draw(n):
        out = np.empty(n, np.uint8)
        for i in range(n):
            out[i] = np.random.randint(0, 256)
        return out

    S = R.vpp_standalone
    H, W = 48, 96
    p = synth.make_pair(40 + wsize, shape=(H, W), hints="random", channels=3, density=0.04, foreground=2)
    l0 = p["left"][..., :C].copy(); r0 = p["right"][..., :C].copy()
    g = (p["hints"] * 0.25).astype(np.float32)
    g_occ = (np.random.default_rng(3).random((H, W)) < 0.2).astype(np.uint8)
    li = l0 if C == 3 else l0[..., 0]; ri = r0 if C == 3 else r0[..., 0]
    kw = dict(wsize=wsize, wsizeAgg_x=16, wsizeAgg_y=3, blending=0.4, use_distance_patch=bool(distance), use_bilateral_patch=bool(bilateral),
              distance_gamma=0.3, bilateral_o_xy=2, bilateral_o_i=3, bilateral_th=.001, uniform_color=bool(uniform), method=method,
              c_occ=0.1)
    nb_seed(21); stn = nb_draw(orc.stream_length(g, wsize, C, uniform)); nb_seed(21)
    la, ra = S.vpp(li, ri, g, g_occ=g_occ.astype(np.float32), **kw)
    lb, rb = orc.vpp(li, ri, g, g_occ=g_occ, stream=stn, mode=1, **kw)
    assert_same(lb, la, "left"); assert_same(rb, ra, "right")
    assert (la != (l0 if C == 3 else l0)).any()
    if bilateral:
        gray = orc.bgr2gray(l0) if C == 3 else l0[..., 0]
        import cv2
        if C == 3:
            assert_same(gray, cv2.cvtColor(l0, cv2.COLOR_BGR2GRAY), "BGR2GRAY")
        want = S._bilateral_filling(g, gray, (wsize - 1) // 2, 2, 3, .001)
        assert want.dtype == np.float64 and np.array_equal(want, want.astype(np.float32))      # float32 values in a float64 array
        assert_same(orc.bilateral_filling(g, gray, (wsize - 1) // 2, 2, 3, .001), want.astype(np.float32), "bilateral filling")


def test_vpp_random_flag_space_vs_reference(orc, R):
    """Seeded random sweep over the signature of vpp() (tests/fuzz_vpp.py): oracle against the reference's numba code with
    numba's own random stream."""
    from numba import njit
    import fuzz_vpp

    @njit
    def nb_seed(s):
        np.random.seed(s)

    @njit
    def nb_draw(n):
        out = np.empty(n, np.uint8)
        for i in range(n):
            out[i] = np.random.randint(0, 256)
        return out

    S = R.vpp_standalone
    for i, li, ri, g, g_occ, stream, kw in fuzz_vpp.cases(40, 77, max_hw=(40, 100)):
        if not (g > 0).any():
            continue
        nb_seed(i); st = nb_draw(stream.size); nb_seed(i)
        la, ra = S.vpp(li, ri, g, g_occ=g_occ.astype(np.float32), **kw)
        lb, rb = orc.vpp(li, ri, g, g_occ=g_occ, stream=st, mode=1, **kw)
        assert_same(lb, la, f"case {i} left {kw}"); assert_same(rb, ra, f"case {i} right {kw}")
