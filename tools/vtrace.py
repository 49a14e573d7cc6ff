"""Debug: per-warp timeline of the v-sweep (library built with -DVPP_TRACE, VPPB200_LIB_SUFFIX=_trace).
Prints, for one CTA and a window of rows, when each warp passed each stage (clock64 cycles relative to the window start)."""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vppstereo_b200 import _lib
from vppstereo_b200.pipeline import VppRsgmPipeline

if "VPPB200_V_RED" in os.environ:
    _lib.set_tuning(_lib.TUNE_SGM_V_RED, int(os.environ["VPPB200_V_RED"]))
B = 64
t = tuple(torch.from_numpy(a).cuda() for a in bench.bench_inputs(0, B))
pipe = VppRsgmPipeline(bench.H, bench.W, 3, batch=B, dmax=192)
for _ in range(2):
    pipe.run_device_serial(*t)
torch.cuda.synchronize()
ROWS, ST = 24, 12
buf = np.zeros(32 * ROWS * ST, np.uint64)
assert _lib.lib().vppb200_debug_vtrace(buf.ctypes.data_as(C.c_void_p), buf.size) == 0
tr = buf.reshape(32, ROWS, ST).astype(np.int64)
nw = int((tr[:, 0, 0] > 0).sum())
t0 = tr[:nw, 0, 0].min()
tr = tr[:nw] - t0
names = ["start", "waited", "deposit", "loop", "blk1", "blk2", "blk3", "blk4", "pushed", "arrived", "prefetch"]
print("row period (cycles):", np.diff(tr[:, :, 0], axis=1).mean(axis=1).round().astype(int).tolist())
print("per-warp mean durations (cycles):  wait | deposit | setup | blk1 | blk2 | blk3 | blk4 | push | arrive | prefetch")
for w in range(nw):
    d = np.diff(tr[w, :, :11], axis=1).mean(axis=0)
    print(f"warp {w:2d} (smsp {w%4}): " + " ".join(f"{x:7.0f}" for x in d) + f"  | row {np.diff(tr[w,:,0]).mean():7.0f}")
print("row 5 timeline (start, waited, loop, loopend, arrived) per warp:")
for w in range(nw):
    r = tr[w, 5]
    print(f"warp {w:2d}: " + " ".join(f"{r[k]-tr[:,5,0].min():7d}" for k in (0, 1, 3, 7, 9)))
np.save(os.path.join("gpurun_out", "vtrace.npy"), tr)
