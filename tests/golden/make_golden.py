#!/usr/bin/env python
"""Generate the committed golden vectors from the UNMODIFIED reference (oracle/_ref, built from /root/reference).

Run where /root/reference exists:   python tests/golden/make_golden.py
Outputs (small, committed): tests/golden/rsgm_*.npz, tests/golden/vpp_cases.npz, tests/golden/vpp_adaptive_cases.npz, tests/golden/occ_cases.npz, tests/golden/rcp_lut.npz

What is pinned
  * rSGM: inputs + the reference's per-stage outputs (census, sha256 of the cost volume and of the aggregated volume,
    WTA / sub-pixel / median disparities, final compute_rsgm output) on crops of the in-repo real pairs
    (thirdparty/stereo-vision/reconstruction/test/*.png) and on a synthetic colour frame.  The reference leaves some
    census pixels unwritten (uninitialised malloc); they are zeroed in its census output before the later stages run
    (oracle/ref.py::zero_unwritten_census) -- that is the parity definition.
  * the RCPSS table of the CPU that produced the sub-pixel values (the instruction is vendor specific).
  * VPP: inputs, pattern stream and outputs of the reference Cython module (`init_rand(seed)`, libc stream) and of the
    numba twin (`vpp()`, generator seeded in-jit) for a set of flag combinations, both methods.
  * occlusion mask: hint maps with foreground boxes and the outputs of the reference's filter.occlusion_heuristic.
"""
import hashlib
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref, oracle as orc  # noqa: E402
from vppstereo_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
TEST_IMGS = "/root/reference/thirdparty/stereo-vision/reconstruction/test/"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def rsgm_case(r, name, left, left_vpp, right_vpp, D, hints=None, valid=None):
    import cv2
    m = r.rsgm
    H, W = left_vpp.shape[:2]
    pad_h, pad_w = (((H // 16) + 1) * 16 - H) % 16, (((W // 16) + 1) * 16 - W) % 16
    p = [pad_w // 2, pad_w - pad_w // 2, pad_h // 2, pad_h - pad_h // 2]
    lp = cv2.copyMakeBorder(left, p[2], p[3], p[0], p[1], cv2.BORDER_REFLECT)
    lvp = cv2.copyMakeBorder(left_vpp, p[2], p[3], p[0], p[1], cv2.BORDER_REFLECT)
    rvp = cv2.copyMakeBorder(right_vpp, p[2], p[3], p[0], p[1], cv2.BORDER_REFLECT)
    ctl, ctr = m._census_transform(lvp, rvp)           # pinned: unwritten pixels zeroed
    dsi = m._hamming_matching(ctl, ctr, D)
    if hints is not None:
        hp = cv2.copyMakeBorder(hints, p[2], p[3], p[0], p[1], cv2.BORDER_CONSTANT, value=0)
        vp = cv2.copyMakeBorder(valid, p[2], p[3], p[0], p[1], cv2.BORDER_CONSTANT, value=0)
        dsi = m._guided_dsi(dsi, hp, vp)
    agg = m._aggregate_dsi(lp, dsi)
    Hp, Wp = ctl.shape
    wl = np.zeros((Hp, Wp), np.float32); r.pyrSGM.matchWTA_SSE(agg, wl, Wp, Hp, D, 0.95)
    sp = wl.copy(); r.pyrSGM.subPixelRefine(agg, sp, Wp, Hp, D, 0)
    sp1 = wl.copy(); r.pyrSGM.subPixelRefine(agg, sp1, Wp, Hp, D, 1)
    med = np.zeros((Hp, Wp), np.float32); r.pyrSGM.median3x3_SSE(sp, med, Wp, Hp)
    wr = np.zeros((Hp, Wp), np.float32); r.pyrSGM.matchWTARight_SSE(agg, wr, Wp, Hp, D, 0.95)
    dl = m._disparity_computation(agg)
    dr = m._right_disparity_computation(agg)
    out_sub = m.compute_rsgm(left, left_vpp, right_vpp, hints=hints, validhints=valid, dmax=D, subpixel=True)
    out_int = m.compute_rsgm(left, left_vpp, right_vpp, hints=hints, validhints=valid, dmax=D, subpixel=False)
    kw = dict(left=left, left_vpp=left_vpp, right_vpp=right_vpp, D=np.int32(D), census_l=ctl, census_r=ctr,
              dsi_sha=np.array(sha(dsi)), agg_sha=np.array(sha(agg)), agg_row=agg[Hp // 2].copy(), wta_l=wl, wta_r=wr,
              subpix0=sp, subpix1=sp1, median=med, disp_l=dl, disp_r=dr, out_sub=out_sub, out_int=out_int)
    if hints is not None:
        kw.update(hints=hints, validhints=valid)
    np.savez_compressed(os.path.join(OUT, f"rsgm_{name}.npz"), **kw)
    print("rsgm", name, left_vpp.shape, "D", D, "mean", float(out_sub.mean()))


def make_rsgm(r):
    import cv2
    L = cv2.imread(TEST_IMGS + "tsukuba_l.png", 0); R = cv2.imread(TEST_IMGS + "tsukuba_r.png", 0)
    rsgm_case(r, "tsukuba_crop", L[100:196, 120:280].copy(), L[100:196, 120:280].copy(), R[100:196, 120:280].copy(), 32)
    L = cv2.imread(TEST_IMGS + "left_000087.png", 0); R = cv2.imread(TEST_IMGS + "right_000087.png", 0)
    # odd crop size -> reflect padding on both axes; D=64 exercises the 2-word lanes, D=72 the partially filled lane
    rsgm_case(r, "kitti_crop", L[200:290, 500:703].copy(), L[200:290, 500:703].copy(), R[200:290, 500:703].copy(), 64)
    rsgm_case(r, "kitti_crop_d72", L[210:280, 400:560].copy(), L[210:280, 400:560].copy(), R[210:280, 400:560].copy(), 72)
    p = synth.make_pair(3, shape=(75, 130), hints="random")
    rsgm_case(r, "synth_colour", p["left"], p["left"], p["right"], 48)
    # guided variant (hints modulate the cost volume, rsgm.py:115-127); P2 guide differs from the matching image
    g = p["hints"]
    vl, vr = orc.vpp(p["left"], p["right"], g, stream=np.arange(100000, dtype=np.uint8), mode=1)
    rsgm_case(r, "synth_guided", p["left"], vl, vr, 48, hints=g, valid=(g > 0).astype(np.float32))


def make_occ(r):
    """filter.occlusion_heuristic (filter.py:246-292) on hint maps with foreground boxes, default and non-default windows."""
    cases = {}
    combos = [((60, 100), "random", 0.08, {}), ((90, 160), "lidar", 0.05, {}), ((48, 64), "random", 0.3, dict(rx=5, ry=9, l=1, g=0.25, th_conf=2)),
              ((70, 120), "random", 0.15, dict(th_filter=1.5)), ((33, 47), "random", 0.5, {})]
    for idx, (shape, kind, density, kw) in enumerate(combos):
        p = synth.make_pair(300 + idx, shape=shape, hints=kind, density=density, foreground=3)
        g = p["hints"].astype(np.float32)
        d, c = r.filter.occlusion_heuristic(g.copy(), **kw)
        full = dict(rx=9, ry=7, l=2, g=0.4375, th_conf=1, th_filter=0.1); full.update(kw)
        k = f"o{idx}_"
        cases.update({k + "g": g, k + "dmap": d, k + "conf": c,
                      k + "params": np.array([full[n] for n in ("rx", "ry", "l", "g", "th_conf", "th_filter")], np.float64)})
        print("occ case", idx, shape, kind, "hints", int((g > 0).sum()), "occluded hints", int((c[g > 0] != 0).sum()), "kept", int((d > 0).sum()))
    np.savez_compressed(os.path.join(OUT, "occ_cases.npz"), **cases)


def make_vpp(r):
    from numba import njit

    @njit
    def nb_seed(s):
        np.random.seed(s)

    @njit
    def nb_draw(n):
        out = np.empty(n, np.uint8)
        for i in range(n):
            out[i] = np.random.randint(0, 256)
        return out

    V, S = r.vpp_core_opt, r.vpp_standalone
    rng = np.random.default_rng(11)
    cases = {}
    combos = [  # C, wsize, direction, uniform, interp, discard, occ_on, dark
        (3, 3, 1, 0, 1, 0, 0, 0), (3, 3, 1, 0, 1, 0, 1, 0), (1, 5, 0, 0, 1, 0, 1, 0), (3, 1, 1, 1, 0, 0, 1, 0),
        (3, 3, 0, 1, 1, 1, 1, 0), (1, 7, 1, 0, 0, 0, 1, 1), (3, 5, 1, 0, 1, 0, 1, 1), (3, 3, 1, 0, 0, 1, 0, 0),
    ]
    H, W = 36, 64
    for idx, (C, wsize, direction, uniform, interp, discard, occ_on, dark) in enumerate(combos):
        p = synth.make_pair(50 + idx, shape=(H, W), hints="random", channels=3, density=0.08)
        l0 = p["left"][..., :C].copy(); r0 = p["right"][..., :C].copy()
        if dark:
            l0 //= 100; r0 //= 100          # values 0..2: lots of zero samples (n_bins book-keeping)
        g = (p["hints"] * 0.2).astype(np.float32)
        g[5, 3] = 7.0; g[6, 60] = 30.5; g[0, 0] = 2.5; g[H - 1, W - 1] = 1.0; g[9, 9] = 3.5
        g_occ = ((rng.random((H, W)) < 0.3) & (occ_on == 1)).astype(np.uint8)
        c, c_occ = 0.4, 0.15
        aggx, aggy = (64, 5) if dark else (16, 3)
        n = orc.stream_length(g, wsize, C, uniform)
        seed = 100 + idx
        st_c = orc.libc_rand_stream(seed, n)
        la, ra = l0.copy(), r0.copy()
        V.init_rand(seed)
        cnt = V.virtual_projection_scan_rnd(la, ra, g, W, H, C, uniform, wsize, direction, c, c_occ, g_occ, discard, interp)
        lm, rm = l0.copy(), r0.copy()
        V.virtual_projection_scan_max_dist(lm, rm, g, W, H, C, uniform, wsize, aggx, aggy, direction, c, c_occ, g_occ, discard, interp)
        nb_seed(seed); st_n = nb_draw(n); nb_seed(seed)
        kw = dict(wsize=wsize, wsizeAgg_x=aggx, wsizeAgg_y=aggy, left2right=bool(direction), blending=c,
                  uniform_color=bool(uniform), c_occ=c_occ, g_occ=g_occ.astype(np.float32), discard_occ=bool(discard),
                  interpolate=bool(interp))
        li = l0 if C == 3 else l0[..., 0]; ri = r0 if C == 3 else r0[..., 0]
        ln, rn = S.vpp(li, ri, g, method="rnd", **kw)
        lnm, rnm = S.vpp(li, ri, g, method="maxDistance", **kw)
        k = f"c{idx}_"
        cases.update({k + "params": np.array([C, wsize, direction, uniform, interp, discard, aggx, aggy, seed, cnt], np.int64),
                      k + "l": l0, k + "r": r0, k + "g": g, k + "g_occ": g_occ, k + "stream_libc": st_c, k + "stream_numba": st_n,
                      k + "cy_rnd_l": la, k + "cy_rnd_r": ra, k + "cy_max_l": lm, k + "cy_max_r": rm,
                      k + "nb_rnd_l": ln, k + "nb_rnd_r": rn, k + "nb_max_l": lnm, k + "nb_max_r": rnm})
        print("vpp case", idx, (C, wsize, direction, uniform, interp, discard, occ_on, dark), "hints", cnt)
    cases["n_cases"] = np.int64(len(combos))
    np.savez_compressed(os.path.join(OUT, "vpp_cases.npz"), **cases)


def make_vpp_adaptive(r):
    """vpp() with the TPAMI adaptive patches (vpp_standalone.py:6-11, :371-394), recorded from the reference's numba code."""
    from numba import njit

    @njit
    def nb_seed(s):
        np.random.seed(s)

    @njit
    def nb_draw(n):
        out = np.empty(n, np.uint8)
        for i in range(n):
            out[i] = np.random.randint(0, 256)
        return out

    S = r.vpp_standalone
    cases = {}
    combos = [  # C, wsize, distance, bilateral, uniform, direction, o_i
        (3, 7, 1, 1, 0, 1, 3), (3, 5, 0, 1, 0, 0, 1), (1, 7, 1, 0, 1, 1, 1), (3, 9, 1, 1, 1, 1, 5), (1, 5, 1, 1, 0, 0, 2),
    ]
    H, W = 44, 80
    for idx, (C, wsize, distance, bilateral, uniform, direction, o_i) in enumerate(combos):
        p = synth.make_pair(300 + idx, shape=(H, W), hints="random", channels=3, density=0.04, foreground=2)
        l0 = p["left"][..., :C].copy(); r0 = p["right"][..., :C].copy()
        g = (p["hints"] * 0.25).astype(np.float32)
        g_occ = (np.random.default_rng(idx).random((H, W)) < 0.2).astype(np.uint8)
        li = l0 if C == 3 else l0[..., 0]; ri = r0 if C == 3 else r0[..., 0]
        n = orc.stream_length(g, wsize, C, uniform)
        kw = dict(wsize=wsize, wsizeAgg_x=16, wsizeAgg_y=3, left2right=bool(direction), blending=0.4, use_distance_patch=bool(distance),
                  use_bilateral_patch=bool(bilateral), distance_gamma=0.3, bilateral_o_xy=2, bilateral_o_i=o_i, bilateral_th=.001,
                  uniform_color=bool(uniform), c_occ=0.1, g_occ=g_occ.astype(np.float32))
        seed = 500 + idx
        nb_seed(seed); st_n = nb_draw(n); nb_seed(seed)
        ln, rn = S.vpp(li, ri, g, method="rnd", **kw)
        lm, rm = S.vpp(li, ri, g, method="maxDistance", **kw)
        k = f"a{idx}_"
        cases.update({k + "params": np.array([C, wsize, distance, bilateral, uniform, direction, o_i], np.int64), k + "l": l0, k + "r": r0,
                      k + "g": g, k + "g_occ": g_occ, k + "stream_numba": st_n, k + "rnd_l": ln, k + "rnd_r": rn, k + "max_l": lm,
                      k + "max_r": rm})
        if bilateral:
            import cv2
            gray = cv2.cvtColor(l0, cv2.COLOR_BGR2GRAY) if C == 3 else l0[..., 0]
            cases[k + "filled"] = S._bilateral_filling(g, gray, (wsize - 1) // 2, 2, o_i, .001).astype(np.float32)
        print("vpp adaptive case", idx, combos[idx], "changed px", int((ln != l0.reshape(ln.shape)).sum()))
    cases["n_cases"] = np.int64(len(combos))
    np.savez_compressed(os.path.join(OUT, "vpp_adaptive_cases.npz"), **cases)


if __name__ == "__main__":
    r = ref.load_pinned()
    parts = sys.argv[1:] or ["rsgm", "vpp", "adaptive", "occ"]          # e.g. `make_golden.py occ` regenerates one family only
    if "rsgm" in parts:
        np.savez_compressed(os.path.join(OUT, "rcp_lut.npz"), lut=orc.rcp_lut())
        make_rsgm(r)
    if "vpp" in parts:
        make_vpp(r)
    if "adaptive" in parts:
        make_vpp_adaptive(r)
    if "occ" in parts:
        make_occ(r)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
