"""CPU: the C oracle against the committed golden vectors recorded from the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle on boxes where /root/reference does not exist."""
import hashlib
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_same, golden_rsgm_names, load_golden_rsgm


@pytest.mark.parametrize("name", golden_rsgm_names())
def test_rsgm_stages(orc, golden_lut, name):
    g = load_golden_rsgm(name)
    D = int(g["D"])
    lv, rv, left = g["left_vpp"], g["right_vpp"], g["left"]
    H, W = lv.shape[:2]
    pad_h, pad_w = (((H // 16) + 1) * 16 - H) % 16, (((W // 16) + 1) * 16 - W) % 16
    p = [pad_w // 2, pad_w - pad_w // 2, pad_h // 2, pad_h - pad_h // 2]
    lvp, rvp, lp = (orc.pad_reflect(a, p[2], p[3], p[0], p[1]) for a in (lv, rv, left))
    gl, gr = (orc.rgb2gray(a) if a.ndim == 3 else a for a in (lvp, rvp))
    Hp, Wp = gl.shape
    cl = np.zeros((Hp, Wp), np.uint32); cr = np.zeros((Hp, Wp), np.uint32)
    orc.census5x5_SSE(gl, cl, Wp, Hp); orc.census5x5_SSE(gr, cr, Wp, Hp)
    assert_same(cl, g["census_l"], "census L"); assert_same(cr, g["census_r"], "census R")
    dsi = np.zeros((Hp, Wp, D), np.uint16); orc.costMeasureCensus5x5_xyd_SSE(cl, cr, dsi, Wp, Hp, D, 1)
    if "hints" in g:
        hp = np.zeros((Hp, Wp), np.float32); vp = np.zeros((Hp, Wp), np.float32)
        hp[p[2]:p[2] + H, p[0]:p[0] + W] = g["hints"]; vp[p[2]:p[2] + H, p[0]:p[0] + W] = g["validhints"]
        dsi = orc.guided_dsi(dsi, hp, vp)
    assert hashlib.sha256(dsi.tobytes()).hexdigest() == str(g["dsi_sha"]), "cost volume sha256"
    agg = np.zeros_like(dsi); orc.aggregate_SSE(lp, dsi, agg, Wp, Hp, D, 11, 17, 0.5, 35)
    assert_same(agg[Hp // 2], g["agg_row"], "aggregated volume, middle row")
    assert hashlib.sha256(agg.tobytes()).hexdigest() == str(g["agg_sha"]), "aggregated volume sha256"
    wl = np.zeros((Hp, Wp), np.float32); orc.matchWTA_SSE(agg, wl, Wp, Hp, D)
    assert_same(wl, g["wta_l"], "WTA left")
    wr = np.zeros((Hp, Wp), np.float32); orc.matchWTARight_SSE(agg, wr, Wp, Hp, D)
    assert_same(wr, g["wta_r"], "WTA right")
    sp = wl.copy(); orc.subPixelRefine(agg, sp, Wp, Hp, D, 0, lut=golden_lut)
    assert_same(sp, g["subpix0"], "sub-pixel (equiangular, recorded RCPSS table)")
    sp1 = wl.copy(); orc.subPixelRefine(agg, sp1, Wp, Hp, D, 1)
    assert_same(sp1, g["subpix1"], "sub-pixel (parabolic)")
    med = np.zeros((Hp, Wp), np.float32); orc.median3x3_SSE(sp, med, Wp, Hp)
    assert_same(med, g["median"], "median")
    dl = med.copy(); orc.linear_interpolate(dl, 15, 3.0); dl = np.clip(dl, 0, None)
    assert_same(dl, g["disp_l"], "left disparity (interp + clip)")


@pytest.mark.parametrize("name", golden_rsgm_names())
def test_compute_rsgm(orc, golden_lut, name):
    g = load_golden_rsgm(name)
    kw = dict(hints=g["hints"], validhints=g["validhints"]) if "hints" in g else {}
    for sub, key in ((True, "out_sub"), (False, "out_int")):
        out = orc.compute_rsgm(g["left"], g["left_vpp"], g["right_vpp"], dmax=int(g["D"]), subpixel=sub,
                               rcp_lut_override=golden_lut, **kw)
        assert_same(out, g[key], f"compute_rsgm subpixel={sub}")


def test_rcp_table_shape(orc, golden_lut):
    """the host table has the structure the kernels rely on: lut[0] = 0, negative, monotone magnitude, ~ -1/(2k)"""
    lut = orc.rcp_lut()
    assert lut.shape == (65536,) and lut[0] == 0.0
    k = np.arange(1, 65536)
    assert np.all(lut[1:] < 0)
    assert np.max(np.abs(lut[1:] * (-2.0 * k) - 1.0)) < 4e-4          # RCPSS: relative error <= 1.5 * 2^-12
    assert golden_lut.shape == (65536,)


def test_vpp_cases(orc, golden_vpp):
    G = golden_vpp
    for i in range(int(G["n_cases"])):
        k = f"c{i}_"
        C, wsize, direction, uniform, interp, discard, aggx, aggy, seed, cnt = [int(v) for v in G[k + "params"]]
        l0, r0, g, g_occ = G[k + "l"], G[k + "r"], G[k + "g"], G[k + "g_occ"]
        H, W = g.shape
        la, ra = l0.copy(), r0.copy()
        n = orc.virtual_projection_scan_rnd(la, ra, g, W, H, C, uniform, wsize, direction, 0.4, 0.15, g_occ, discard, interp,
                                            stream=G[k + "stream_libc"], mode=0)
        assert n == cnt
        assert_same(la, G[k + "cy_rnd_l"], f"case {i} cython rnd L"); assert_same(ra, G[k + "cy_rnd_r"], f"case {i} cython rnd R")
        la, ra = l0.copy(), r0.copy()
        orc.virtual_projection_scan_max_dist(la, ra, g, W, H, C, uniform, wsize, aggx, aggy, direction, 0.4, 0.15, g_occ, discard,
                                             interp, mode=0)
        assert_same(la, G[k + "cy_max_l"], f"case {i} cython maxDistance L"); assert_same(ra, G[k + "cy_max_r"], f"case {i} cython maxDistance R")
        kw = dict(wsize=wsize, wsizeAgg_x=aggx, wsizeAgg_y=aggy, left2right=bool(direction), blending=0.4, uniform_color=bool(uniform),
                  c_occ=0.15, g_occ=g_occ, discard_occ=bool(discard), interpolate=bool(interp))
        li = l0 if C == 3 else l0[..., 0]; ri = r0 if C == 3 else r0[..., 0]
        ln, rn = orc.vpp(li, ri, g, method="rnd", stream=G[k + "stream_numba"], mode=1, **kw)
        assert_same(ln, G[k + "nb_rnd_l"], f"case {i} numba rnd L"); assert_same(rn, G[k + "nb_rnd_r"], f"case {i} numba rnd R")
        ln, rn = orc.vpp(li, ri, g, method="maxDistance", mode=1, **kw)
        assert_same(ln, G[k + "nb_max_l"], f"case {i} numba maxDistance L"); assert_same(rn, G[k + "nb_max_r"], f"case {i} numba maxDistance R")


def _adaptive_cases():
    G = dict(np.load(os.path.join(GOLDEN, "vpp_adaptive_cases.npz")))
    for i in range(int(G["n_cases"])):
        k = f"a{i}_"
        C, wsize, distance, bilateral, uniform, direction, o_i = [int(v) for v in G[k + "params"]]
        l0, r0 = G[k + "l"], G[k + "r"]
        kw = dict(wsize=wsize, wsizeAgg_x=16, wsizeAgg_y=3, left2right=bool(direction), blending=0.4, use_distance_patch=bool(distance),
                  use_bilateral_patch=bool(bilateral), distance_gamma=0.3, bilateral_o_xy=2, bilateral_o_i=o_i, bilateral_th=.001,
                  uniform_color=bool(uniform), c_occ=0.1, g_occ=G[k + "g_occ"])
        yield i, G, k, (l0 if C == 3 else l0[..., 0]), (r0 if C == 3 else r0[..., 0]), kw


def test_vpp_adaptive_cases(orc):
    """vpp() with distance-based / bilateral adaptive patches: oracle against outputs recorded from the reference's numba code."""
    for i, G, k, li, ri, kw in _adaptive_cases():
        ln, rn = orc.vpp(li, ri, G[k + "g"], method="rnd", stream=G[k + "stream_numba"], mode=1, **kw)
        assert_same(ln, G[k + "rnd_l"], f"adaptive case {i} rnd L"); assert_same(rn, G[k + "rnd_r"], f"adaptive case {i} rnd R")
        lm, rm = orc.vpp(li, ri, G[k + "g"], method="maxDistance", mode=1, **kw)
        assert_same(lm, G[k + "max_l"], f"adaptive case {i} maxDistance L"); assert_same(rm, G[k + "max_r"], f"adaptive case {i} maxDistance R")
        if k + "filled" in G:
            gray = orc.bgr2gray(G[k + "l"]) if G[k + "l"].shape[-1] == 3 else np.ascontiguousarray(G[k + "l"][..., 0])
            assert_same(orc.bilateral_filling(G[k + "g"], gray, (kw["wsize"] - 1) // 2, 2, kw["bilateral_o_i"], .001), G[k + "filled"],
                        f"adaptive case {i} bilateral filling")


def test_libc_stream_matches_golden(golden_vpp):
    """the product's glibc rand() restatement (C-ABI host helper) regenerates the recorded libc streams"""
    from vppstereo_b200 import vpp_core_opt as core
    G = golden_vpp
    for i in range(int(G["n_cases"])):
        seed = int(G[f"c{i}_params"][8])
        want = G[f"c{i}_stream_libc"]
        core.init_rand(seed)
        assert_same(core.draw_pattern(want.size), want, f"libc stream seed {seed}")


def test_occlusion_heuristic_golden(orc):
    """filter.occlusion_heuristic (filter.py:246-292): oracle against the outputs recorded from the reference's numba code."""
    g = dict(np.load(os.path.join(GOLDEN, "occ_cases.npz")))
    n = len([k for k in g if k.endswith("_g")])
    assert n >= 5
    for i in range(n):
        k = f"o{i}_"
        rx, ry, l, gg, thc, thf = g[k + "params"]
        d, c = orc.occlusion_heuristic(g[k + "g"], rx=int(rx), ry=int(ry), l=l, g=gg, th_conf=thc, th_filter=thf)
        assert_same(c, g[k + "conf"], f"occlusion mask, case {i}")
        assert_same(d, g[k + "dmap"], f"filtered hints, case {i}")
