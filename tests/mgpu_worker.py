"""Worker of tests/test_gpu_multi.py (one process per GPU, launched by torch.distributed.run): frame-sharded VPP + rSGM on real
kernels, disparities gathered with dist.PeerGather (copy engines over peer-mapped memory) and with the NCCL all_gather;
every rank then recomputes ALL ranks' shards alone on its own GPU and compares bit for bit (SURVEY.md 8e row 1 / 7.2 last row:
gathered == single-GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    outdir = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from vppstereo_b200 import dist as vd, synth
    from vppstereo_b200.pipeline import VppRsgmPipeline
    H, W, D, B, K = 72, 160, 64, 3, 5

    def shard_inputs(r, k):
        fr = [synth.make_pair(500 + 100 * r + 10 * k + i, shape=(H, W), hints="lidar") for i in range(B)]
        return tuple(torch.from_numpy(np.stack([p[key] for p in fr])).to(dev) for key in ("left", "right", "hints"))

    msgs = []
    for mode in ("p2p", "nccl"):
        pipe = VppRsgmPipeline(H, W, 3, batch=B, dmax=D, device=dev, seed=100 + rank)
        pg = vd.PeerGather((B, H, W), torch.float32, dev, depth=2)
        if mode == "nccl":
            pg.available = False                       # the fallback: one NCCL all_gather per step
        elif not pg.available:
            msgs.append(f"p2p unavailable: {pg.why}")
        got = []
        for k in range(K):                             # K > depth: the slots are reused, the credits flow
            out = pipe.run_device(*shard_inputs(rank, k))
            pg.push(k, out)
            full = pg.wait(k)
            got.append(full.clone())
            pg.release(k)
        torch.cuda.synchronize(dev)
        dist.barrier()
        # every rank alone: all shards of all steps with the owners' seeds
        for r in range(world):
            solo = VppRsgmPipeline(H, W, 3, batch=B, dmax=D, device=dev, seed=100 + r)
            for k in range(K):
                want = solo.run_device(*shard_inputs(r, k)).clone()
                torch.cuda.synchronize(dev)
                a, b = got[k][r].cpu().numpy().view(np.uint32), want.cpu().numpy().view(np.uint32)
                if not np.array_equal(a, b):
                    msgs.append(f"{mode}: step {k} shard {r}: {(a != b).sum()} of {a.size} values differ")
            solo.close()
        pipe.close()
        msgs.append(f"{mode}: available={pg.available} ({pg.why})")
        pg.close()
    bad = [m for m in msgs if "differ" in m]
    with open(os.path.join(outdir, f"rank{rank}.txt"), "w") as f:
        f.write(("FAIL\n" if bad else "OK\n") + "\n".join(msgs) + "\n")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
