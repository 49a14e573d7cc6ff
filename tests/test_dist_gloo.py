"""CPU, world_size 2 over gloo: the N>1 host path -- contiguous frame shards, uneven shard sizes, gather to all ranks and
to one rank -- reproduces the single-process result.  (The per-frame compute is a deterministic stand-in: kernels need a
GPU; this test covers the sharding / collective logic the 8-GPU run relies on.)"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, n_frames, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from vppstereo_b200 import dist as vd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        def make_inputs(lo, hi):
            return torch.arange(lo, hi, dtype=torch.float32)

        def process(idx):                       # stand-in for vpp + compute_rsgm: a frame-local function of the frame id
            return (idx[:, None, None] * 3.0 + torch.arange(6, dtype=torch.float32).reshape(2, 3)).contiguous()

        full = vd.run_sharded(n_frames, make_inputs, process, chunk=2)
        lo, hi = vd.shard_range(n_frames, rank, world)
        root_only = vd.gather_frames(process(make_inputs(lo, hi)), n_frames, dst=0)
        q.put((rank, full.numpy(), None if root_only is None else root_only.numpy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [8, 7])
def test_two_rank_gather(n_frames):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() + n_frames) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, n_frames, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = (np.arange(n_frames, dtype=np.float32)[:, None, None] * 3.0 + np.arange(6, dtype=np.float32).reshape(2, 3))
    for rank, full, root_only in results:
        assert np.array_equal(full, want), f"rank {rank} all_gather result"
        if rank == 0:
            assert np.array_equal(root_only, want)
        else:
            assert root_only is None
