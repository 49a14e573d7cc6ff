"""2-GPU probe (torch.distributed.run): how long does one PeerGather push take, alone and beside a compute kernel?"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vppstereo_b200.dist import PeerGather
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
shape = (64, 375, 1242)
pg = PeerGather(shape, torch.float32, dev, depth=2)
x = torch.randn(shape, device=dev)
cs = pg.copy_stream
def t(fn, n=20):
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(cs); 
    for _ in range(n): fn()
    e1.record(cs); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
peer = (rank + 1) % world
r = {}
r["bulk peer copy 119MB"] = t(lambda: pg._copy(cs, pg.peer_bufs[peer][0, rank], x))
r["bulk self copy 119MB"] = t(lambda: pg._copy(cs, pg.bufs[0, rank], x))
r["4B peer copy"] = t(lambda: pg._copy(cs, pg.peer_words[peer][0, rank:rank+1], pg.ticks[5:6]))
with torch.cuda.stream(cs):
    r["torch self copy_ 119MB"] = t(lambda: pg.bufs[1, rank].copy_(x, non_blocking=True))
# full push/wait/release cycles
k0 = [0]
def cyc():
    k = k0[0]; k0[0] += 1
    pg.push(k, x); pg.wait(k); pg.release(k)
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter()
for _ in range(50): cyc()
torch.cuda.synchronize()
r["push+wait+release cycle (wall)"] = (time.perf_counter() - t0) / 50 * 1e3
if rank == 0:
    for a, b in r.items(): print(f"{a:40s} {b:8.3f} ms")
pg.close(); dist.destroy_process_group()
