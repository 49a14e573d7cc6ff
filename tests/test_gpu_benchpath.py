"""Parity of the code path bench.py times (BASELINE configs[1]): batch 64 at the KITTI shape through
VppRsgmPipeline.run_device -- the device-generated pattern, the three-stream software pipeline, and the v-sweep in the shape the
planner picks for that batch (wide strips, NS >= 131, teams that process several frames back to back with the halo tags carried
across frames).  Reference: models/rsgm/rsgm.py:250-294 on the same frames, vpp_standalone.py:243-369 for the projection.

Three layers: (a) the full-size run against the oracle on frames taken from every round of the teams, (b) the same kernel
instantiations forced onto small frames (seconds), (c) the hash bench.py asserts (tests/golden/bench_check.json)."""
import concurrent.futures as cf
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, assert_same

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    assert torch.cuda.is_available()
    from vppstereo_b200 import _lib, rsgm, synth, vpp_core_opt
    _lib.lib()
    return _lib, rsgm, synth, vpp_core_opt


def _oracle_frame(args):
    """worker: VPP (numba arithmetic, the device generator's stream restated on the host) + compute_rsgm by the oracle"""
    left, right, hints, seed, f, D, lut = args
    from oracle import oracle as orc
    from vppstereo_b200 import vpp_core_opt
    pattern = vpp_core_opt.device_pattern(seed, f, orc.stream_length(hints, 3, left.shape[2], False))
    lw, rw = orc.vpp(left, right, hints, stream=pattern, mode=1)
    return lw, rw, orc.compute_rsgm(left, lw, rw, dmax=D, rcp_lut_override=lut)


def test_device_pattern_restatement(mods):
    """vpp_core_opt.device_pattern == what the kernels draw: one frame projected with the device generator equals the same
    frame projected with the restated stream passed in explicitly."""
    import torch
    _lib, _, synth, core = mods
    from vppstereo_b200 import vpp_standalone
    p = synth.make_pair(3, shape=(60, 150), hints="random")
    L, R, G = (torch.from_numpy(p[k]).cuda() for k in ("left", "right", "hints"))
    n = int(core.draws_per_frame(p["hints"], 3, 3, False).sum())
    for seed in (0, 7, 2**63 + 12345):
        a = vpp_standalone.vpp(L, R, G, seed=seed)                       # vpp() sets bit 63 of a given seed
        b = vpp_standalone.vpp(L, R, G, pattern=core.device_pattern(seed | (1 << 63), 0, n))
        assert_same(a[0].cpu().numpy(), b[0].cpu().numpy(), f"left, seed {seed}")
        assert_same(a[1].cpu().numpy(), b[1].cpu().numpy(), f"right, seed {seed}")


def test_pipeline_batch64_kitti_vs_oracle(mods, orc, golden_lut):
    """(a) The benchmarked configuration itself: 64 K-shape frames per call, two calls back to back (so both buffer sets and the
    cross-call overlap are in play), frames from every round of the v-sweep's teams compared bit for bit with the oracle: the
    projected pair AND the final disparities."""
    import torch
    from vppstereo_b200.pipeline import VppRsgmPipeline
    _lib, _, synth, _ = mods
    _lib.set_tuning(_lib.TUNE_RCP_HOST, 0)
    try:
        B, D = 64, 192
        uniq = [synth.make_pair(f, shape="K", hints="lidar") for f in range(16)]
        idx = [(5 * i) % 16 for i in range(B)]
        left = torch.from_numpy(np.stack([uniq[i]["left"] for i in idx])).cuda()
        right = torch.from_numpy(np.stack([uniq[i]["right"] for i in idx])).cuda()
        hints = torch.from_numpy(np.stack([uniq[i]["hints"] for i in idx])).cuda()
        pipe = VppRsgmPipeline(375, 1242, 3, batch=B, dmax=D, seed=77)
        o1 = torch.empty((B, 375, 1242), dtype=torch.float32, device="cuda"); o2 = torch.empty_like(o1)
        pipe.run_device(left, right, hints, out=o1, inputs_ready=True)
        pipe.run_device(left.flip(0).contiguous(), right.flip(0).contiguous(), hints.flip(0).contiguous(), out=o2, inputs_ready=True)
        torch.cuda.synchronize()
        lv2, rv2 = pipe.lv.cpu().numpy(), pipe.rv.cpu().numpy()          # projected pair of the latest call
        got1, got2 = o1.cpu().numpy(), o2.cpu().numpy()
        checks = [(1, f) for f in (0, 17, 18, 35, 54, 63)] + [(2, f) for f in (0, 36, 63)]
        jobs = []
        for call, f in checks:
            src = idx[f] if call == 1 else idx[B - 1 - f]
            jobs.append((uniq[src]["left"], uniq[src]["right"], uniq[src]["hints"], pipe.pattern_seed(call), f, D, golden_lut))
        # threads: the oracle is C behind ctypes (the GIL is released during the calls, no shared mutable state)
        with cf.ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            res = list(ex.map(_oracle_frame, jobs))
        for (call, f), (lw, rw, want) in zip(checks, res):
            if call == 2:
                assert_same(lv2[f], lw, f"call 2 frame {f}: projected left"); assert_same(rv2[f], rw, f"call 2 frame {f}: projected right")
            assert_same((got1 if call == 1 else got2)[f], want, f"call {call} frame {f}: disparities")
        pipe.close()
    finally:
        _lib.set_tuning(_lib.TUNE_RCP_HOST, 1)


@pytest.fixture
def tuning(mods):
    _lib = mods[0]
    yield lambda key, value: _lib.set_tuning(key, value)
    _lib.set_tuning(_lib.TUNE_SGM_MAX_STRIP, 0)
    _lib.set_tuning(_lib.TUNE_SGM_CLUSTERS, 0)


@pytest.mark.parametrize("teams", [1, 2])
@pytest.mark.parametrize("strip", [128, 160, 192])
@pytest.mark.parametrize("shape,D,n", [((22, 330), 48, 5), ((20, 330), 64, 7), ((14, 400), 192, 5), ((18, 500), 96, 6), ((16, 290), 8, 7)])
def test_wide_strips_multi_round(mods, orc, tuning, strip, teams, shape, D, n):
    """(b) The v-sweep instantiations of the benchmark (4-6 column groups per CTA: NS = 131 / 163 / 195, FULL and guarded) with
    fewer teams than frames, so that every team walks several frames back to back and the halo tags, the row parity and the
    operand prefetch run on across the frame boundary.  Small frames, all frames against the oracle."""
    _lib, rsgm, synth, _ = mods
    tuning(_lib.TUNE_SGM_MAX_STRIP, strip)
    tuning(_lib.TUNE_SGM_CLUSTERS, teams)
    frames = [synth.make_pair(900 + D + f, shape=shape, hints="random") for f in range(n)]
    left = np.stack([p["left"] for p in frames]); right = np.stack([p["right"] for p in frames])
    got = rsgm.compute_rsgm(left, left, right, dmax=D)
    for f, p in enumerate(frames):
        want = orc.compute_rsgm(p["left"], p["left"], p["right"], dmax=D)
        assert_same(got[f], want, f"strip={strip} teams={teams} {shape} D={D} frame {f}/{n}")


def test_bench_check_hash(mods):
    """(c) bench.py asserts sha256 hashes of the projected pair and of the disparities of frames 0 and 63 of one extra batch
    against tests/golden/bench_check.json (made by tests/golden/make_bench_check.py with the oracle).  Here: the same probe on
    the same inputs through the same function."""
    import bench
    with open(os.path.join(GOLDEN, "bench_check.json")) as f:
        want = json.load(f)
    got = bench.parity_probe(batch=64)
    assert got["sha256"] == want["sha256"], f"bench parity probe: {got['sha256']} vs committed {want['sha256']}"
