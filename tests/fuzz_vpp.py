"""Randomised parity sweep of the VPP kernels against the oracle: case generator for the tests, and a stand-alone runner
(on a GPU box: python tests/fuzz_vpp.py [n] [seed]).  Test infrastructure: imports oracle/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vppstereo_b200 import synth
from oracle import oracle as orc

def cases(n_cases, seed, max_hw=(70, 400)):
    """Yields (index, left, right, hints, g_occ, pattern, kwargs of vpp()) over the whole flag space of vpp()."""
    rng = np.random.default_rng(seed)
    for i in range(n_cases):
        H, W = int(rng.integers(8, max_hw[0])), int(rng.integers(16, max_hw[1]))
        C = int(rng.choice([1, 3]))
        wsize = int(rng.choice([1, 3, 5, 7, 9]))
        agg = (int(rng.choice([1, 3, 8, 16, 33, 64])), int(rng.choice([1, 3, 5, 7])))
        method = str(rng.choice(["rnd", "maxDistance"]))
        density = float(rng.choice([0.01, 0.05, 0.2, 0.7]))
        scale = float(rng.choice([0.05, 0.3, 1.0, 3.0]))          # up to disparities beyond the image width
        kw = dict(wsize=wsize, wsizeAgg_x=agg[0], wsizeAgg_y=agg[1], left2right=bool(rng.integers(2)), blending=float(rng.choice([0.4, 0.9, 1.0])),
                  use_distance_patch=bool(rng.integers(2)), use_bilateral_patch=bool(rng.integers(2)), distance_gamma=float(rng.choice([0.3, 1.0])),
                  bilateral_o_xy=int(rng.choice([1, 2, 3])), bilateral_o_i=int(rng.choice([1, 3, 10])), bilateral_th=float(rng.choice([.001, .2])),
                  uniform_color=bool(rng.integers(2)), method=method, c_occ=float(rng.choice([0.0, 0.1, 0.5])), discard_occ=bool(rng.integers(2)),
                  interpolate=bool(rng.integers(2)))
        p = synth.make_pair(int(rng.integers(1000)), shape=(H, W), hints="random", channels=3, density=density, foreground=int(rng.integers(3)))
        l0 = np.ascontiguousarray(p["left"][..., :C]); r0 = np.ascontiguousarray(p["right"][..., :C])
        if rng.integers(4) == 0:
            l0 //= 100; r0 //= 100
        g = (p["hints"] * scale).astype(np.float32)
        g_occ = (rng.random((H, W)) < float(rng.choice([0.0, 0.3]))).astype(np.uint8)
        li = l0 if C == 3 else l0[..., 0]; ri = r0 if C == 3 else r0[..., 0]
        pos = g[g > 0]
        if kw["use_distance_patch"] and (pos.size == 0 or pos.min() == pos.max()):
            kw["use_distance_patch"] = False
        stream = rng.integers(0, 256, max(orc.stream_length(g, wsize, C, kw["uniform_color"]), 1), dtype=np.uint8)
        yield i, li, ri, g, g_occ, stream, kw


def main(n_cases, seed):
    from vppstereo_b200 import vpp_standalone as vs
    bad = 0
    for i, li, ri, g, g_occ, stream, kw in cases(n_cases, seed):
        lw, rw = orc.vpp(li, ri, g, g_occ=g_occ, stream=stream, mode=1, **kw)
        lg, rg = vs.vpp(li, ri, g, g_occ=g_occ, pattern=stream, **kw)
        ok = np.array_equal(lg, lw) and np.array_equal(rg, rw)
        if not ok:
            bad += 1
            print("MISMATCH case", i, li.shape, kw, "diff px", int((lg != lw).sum()), int((rg != rw).sum()), flush=True)
    print(f"fuzz: {n_cases - bad}/{n_cases} cases bit-exact (seed {seed})")
    return bad

if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]) if len(sys.argv) > 1 else 150, int(sys.argv[2]) if len(sys.argv) > 2 else 0) else 0)
