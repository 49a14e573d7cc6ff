"""GPU parity tests of the occlusion heuristic (filter.py:246-292 -> csrc/filter.cu): CUDA kernels against the CPU oracle
and against golden vectors recorded from the reference's numba code; bit-exact (mask uint8, hints float32 bit patterns)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_same

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    assert torch.cuda.is_available()
    from vppstereo_b200 import filter as flt, vpp_standalone, synth, _lib
    _lib.lib()
    return flt, vpp_standalone, synth


def test_golden_occlusion(mods):
    flt, _, _ = mods
    g = dict(np.load(os.path.join(GOLDEN, "occ_cases.npz")))
    n = len([k for k in g if k.endswith("_g")])
    for i in range(n):
        k = f"o{i}_"
        rx, ry, l, gg, thc, thf = g[k + "params"]
        d, c = flt.occlusion_heuristic(g[k + "g"], rx=int(rx), ry=int(ry), l=l, g=gg, th_conf=thc, th_filter=thf)
        assert c.dtype == np.uint8 and d.dtype == np.float32
        assert_same(c, g[k + "conf"], f"occlusion mask, case {i}")
        assert_same(d, g[k + "dmap"], f"filtered hints, case {i}")


@pytest.mark.parametrize("shape,kind,density,fg,kw", [
    ((1, 1), "random", 1.0, 0, {}), ((1, 40), "random", 0.5, 1, {}), ((40, 1), "random", 0.5, 0, {}), ((37, 53), "random", 0.3, 2, {}),
    ((96, 200), "lidar", 0.05, 3, {}), ((120, 160), "random", 0.05, 4, dict(rx=5, ry=11, l=1, g=0.3, th_conf=2)),
    ((64, 64), "random", 0.6, 2, dict(rx=3, ry=3)), ((50, 90), "random", 0.2, 3, dict(th_filter=2)),
    ((480, 640), "random", 0.05, 6, {}), ((375, 1242), "lidar", 0.05, 8, {}), ((1988, 2880), "random", 0.05, 12, {}),
])
def test_occlusion_vs_oracle(mods, orc, shape, kind, density, fg, kw):
    flt, _, synth = mods
    if shape[0] * shape[1] > 10 ** 6:          # Middlebury size: cheap hint map (the synthetic images are not needed here)
        rng = np.random.default_rng(5)
        yy, xx = np.mgrid[0:shape[0], 0:shape[1]]
        dgt = (20.0 + 200.0 * yy / shape[0] + 9.0 * np.sin(xx / 131.0)).astype(np.float32)
        for _ in range(fg):
            by, bx = int(rng.integers(0, shape[0] - 300)), int(rng.integers(0, shape[1] - 400))
            dgt[by:by + 300, bx:bx + 400] += np.float32(rng.uniform(20, 60))
        g = np.where(rng.random(shape) < density, dgt + rng.normal(0, 0.25, shape), 0).astype(np.float32)
    else:
        g = synth.make_pair(11, shape=shape, hints=kind, density=density, foreground=fg)["hints"].astype(np.float32)
    dw, cw = orc.occlusion_heuristic(g, **kw)
    d, c = flt.occlusion_heuristic(g, **kw)
    assert_same(c, cw, "occlusion mask"); assert_same(d, dw, "filtered hints")


def test_empty_and_full_maps(mods, orc):
    flt, _, _ = mods
    z = np.zeros((20, 30), np.float32)
    d, c = flt.occlusion_heuristic(z)
    assert not d.any() and (c == 1).all()
    f = np.full((20, 30), 3.0, np.float32)
    dw, cw = orc.occlusion_heuristic(f)
    d, c = flt.occlusion_heuristic(f)
    assert_same(c, cw, "mask"); assert_same(d, dw, "hints")


def test_batch_on_device_feeds_vpp(mods, orc):
    """[N,H,W] CUDA tensor in -> CUDA tensors out; the mask goes into vpp(g_occ=...) without a host hop (test.py:154-176)."""
    import torch
    flt, vs, synth = mods
    frames = [synth.make_pair(40 + i, shape=(72, 128), hints="random", density=0.08, foreground=3) for i in range(3)]
    g = np.stack([f["hints"] for f in frames]).astype(np.float32)
    gd = torch.from_numpy(g).cuda()
    d, c = flt.occlusion_heuristic(gd)
    assert d.is_cuda and c.is_cuda and c.dtype == torch.uint8 and tuple(c.shape) == g.shape
    for i, f in enumerate(frames):
        dw, cw = orc.occlusion_heuristic(g[i])
        assert_same(c[i].cpu().numpy(), cw, f"mask {i}"); assert_same(d[i].cpu().numpy(), dw, f"hints {i}")
        assert int((cw[g[i] > 0] != 0).sum()) > 0, "the case must contain occluded hints"
        n = orc.stream_length(g[i], 3, 3, False)
        stream = np.random.default_rng(i).integers(0, 256, n, dtype=np.uint8)
        for method in ("rnd", "maxDistance"):
            lw, rw = orc.vpp(f["left"], f["right"], g[i], method=method, c_occ=0.1, g_occ=cw, stream=stream, mode=1, wsizeAgg_x=16)
            lg, rg = vs.vpp(torch.from_numpy(f["left"]).cuda(), torch.from_numpy(f["right"]).cuda(), gd[i], method=method,
                            c_occ=0.1, g_occ=c[i], pattern=stream, wsizeAgg_x=16)
            assert_same(lg.cpu().numpy(), lw, f"left {method} {i}"); assert_same(rg.cpu().numpy(), rw, f"right {method} {i}")
