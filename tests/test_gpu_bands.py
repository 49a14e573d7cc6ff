"""Row-band split of an oversized frame: the stages that shard exactly by rows -- VPP rnd (hint rows within the
patch radius), census (3-row halo: 2 for the window + 1 for the flat-stream wrap of the edge columns) and the Hamming cost
volume (row local) -- give, band by band, exactly what the whole frame gives.  The NCCL path over two ranks runs when the box
has two GPUs (the band arithmetic itself is also covered on CPU by tests/test_dist_gloo.py)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_same

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import torch
    assert torch.cuda.is_available()
    from vppstereo_b200 import vpp_core_opt, vpp_standalone, synth, pyrSGM, dist, _lib
    _lib.lib()
    return vpp_core_opt, vpp_standalone, synth, pyrSGM, dist


@pytest.mark.parametrize("wsize,world,occ_frac,discard", [(3, 2, 0.0, 0), (5, 3, 0.3, 0), (7, 4, 0.3, 1), (1, 5, 0.0, 0)])
def test_vpp_rnd_bands_union_equals_full_scan(mods, orc, wsize, world, occ_frac, discard):
    import torch
    core, standalone, synth, _, vd = mods
    H, W, C = 61, 300, 3
    p = synth.make_pair(9 + wsize, shape=(H, W), hints="random", density=0.08)
    g = (p["hints"] * 0.3).astype(np.float32)
    g_occ = (np.random.default_rng(1).random((H, W)) < occ_frac).astype(np.uint8)
    pattern = np.random.default_rng(wsize).integers(0, 256, orc.stream_length(g, wsize, C, 0), dtype=np.uint8)
    lw, rw = orc.vpp(p["left"], p["right"], g, wsize=wsize, g_occ=g_occ, c_occ=0.2, discard_occ=bool(discard), stream=pattern, mode=1)
    lt, rt = torch.from_numpy(p["left"]).cuda(), torch.from_numpy(p["right"]).cuda()
    gt, ot = torch.from_numpy(g).cuda(), torch.from_numpy(g_occ).cuda()
    for rank in reversed(range(world)):               # any order: the bands are independent
        lo, hi = vd.shard_range(H, rank, world)
        core._scan("rnd", lt, rt, gt, W, H, C, 0, wsize, None, 1, 0.4, 0.2, ot, discard, 1, pattern, 1, want_counts=False, rows=(lo, hi))
        if rank == world - 1 and world > 1:           # rows of the other bands are still the input
            assert_same(lt[:lo].cpu().numpy(), p["left"][:lo], "rows outside the band are untouched")
    assert_same(lt.cpu().numpy(), lw, "left image, union of bands"); assert_same(rt.cpu().numpy(), rw, "right image, union of bands")
    # single-process front-end = whole frame
    lb, rb = vd.vpp_rnd_banded(lt * 0 + torch.from_numpy(p["left"]).cuda(), torch.from_numpy(p["right"]).cuda(), gt, pattern=pattern,
                               wsize=wsize, g_occ=ot, c_occ=0.2, discard_occ=bool(discard))
    assert_same(lb.cpu().numpy(), lw, "vpp_rnd_banded (world 1) left"); assert_same(rb.cpu().numpy(), rw, "vpp_rnd_banded (world 1) right")
    with pytest.raises(ValueError):
        core._scan("max_dist", lt, rt, gt, W, H, C, 0, wsize, (16, 3), 1, 0.4, 0.2, ot, discard, 1, None, 1, want_counts=False, rows=(0, 8))


@pytest.mark.parametrize("world", [2, 3, 5])
def test_census_and_cost_bands(mods, orc, world):
    """census on (band + 3 halo rows) then cropped == census of the whole frame on the band's rows, except where the whole
    frame itself leaves rows unwritten (first / last two rows); the cost volume of a band is that band of the cost volume."""
    _, _, synth, pyrSGM, vd = mods
    H, W, D = 96, 128, 32
    rng = np.random.default_rng(world)
    left = rng.integers(0, 256, (H, W), dtype=np.uint8); right = rng.integers(0, 256, (H, W), dtype=np.uint8)
    cl = np.zeros((H, W), np.uint32); cr = np.zeros((H, W), np.uint32)
    pyrSGM.census5x5_SSE(left, cl, W, H); pyrSGM.census5x5_SSE(right, cr, W, H)
    dsi = np.zeros((H, W, D), np.uint16); pyrSGM.costMeasureCensus5x5_xyd_SSE(cl, cr, dsi, W, H, D, 1)
    for rank in range(world):
        (lo, hi), (rlo, rhi) = vd.band_with_halo(H, rank, world, 3)
        sub = np.ascontiguousarray(left[rlo:rhi]); got = np.zeros(sub.shape, np.uint32)
        pyrSGM.census5x5_SSE(sub, got, W, rhi - rlo)
        a, b = max(lo, 2), min(hi, H - 3)             # rows the whole-frame census defines with full windows
        assert_same(got[a - rlo:b - rlo], cl[a:b], f"census band {rank}/{world}")
        # cost volume: row local given the census rows (rows 0,1,H-2,H-1 of the FRAME are the constant 12: compare inner rows)
        bl, br = np.ascontiguousarray(cl[rlo:rhi]), np.ascontiguousarray(cr[rlo:rhi])
        bd = np.zeros((rhi - rlo, W, D), np.uint16); pyrSGM.costMeasureCensus5x5_xyd_SSE(bl, br, bd, W, rhi - rlo, D, 1)
        a2, b2 = max(lo, rlo + 2, 2), min(hi, rhi - 2, H - 2)
        assert_same(bd[a2 - rlo:b2 - rlo], dsi[a2:b2], f"cost volume band {rank}/{world}")


def _nccl_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from vppstereo_b200 import dist as vd, synth
        p = synth.make_pair(5, shape=(90, 260), hints="random", density=0.06)
        pattern = np.random.default_rng(0).integers(0, 256, 200000, dtype=np.uint8)
        lc, rc = vd.vpp_rnd_banded(torch.from_numpy(p["left"]).cuda(), torch.from_numpy(p["right"]).cuda(),
                                   torch.from_numpy(p["hints"]).cuda(), pattern=pattern, wsize=5)
        q.put((rank, lc.cpu().numpy(), rc.cpu().numpy()))
    finally:
        dist.destroy_process_group()


def test_vpp_rnd_banded_two_gpus(orc):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from vppstereo_b200 import synth
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_nccl_worker, args=(r, world, 29650 + os.getpid() % 300, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = synth.make_pair(5, shape=(90, 260), hints="random", density=0.06)
    pattern = np.random.default_rng(0).integers(0, 256, 200000, dtype=np.uint8)
    lw, rw = orc.vpp(p["left"], p["right"], p["hints"], wsize=5, stream=pattern, mode=1)
    for rank, lc, rc in res:
        assert_same(lc, lw, f"rank {rank} left"); assert_same(rc, rw, f"rank {rank} right")
