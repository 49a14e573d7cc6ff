"""Experiment: v-sweep strip width / frames in flight (vppb200_set_tuning) on the benchmark batch."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from vppstereo_b200 import _lib, synth
from vppstereo_b200.pipeline import VppRsgmPipeline
import ctypes
dev = torch.device('cuda', 0)
B = 64
frames = [synth.make_pair(f, shape="K", hints="lidar") for f in range(4)]
idx = [i % 4 for i in range(B)]
left = torch.from_numpy(np.stack([frames[i]["left"] for i in idx])).to(dev)
right = torch.from_numpy(np.stack([frames[i]["right"] for i in idx])).to(dev)
hints = torch.from_numpy(np.stack([frames[i]["hints"] for i in idx])).to(dev)
pipe = VppRsgmPipeline(375, 1242, 3, batch=B, dmax=192, device=dev)
L = _lib.lib()
for strip in [int(a) for a in sys.argv[1:]] or [0, 192, 160, 128, 96, 64]:
    _lib.set_tuning(_lib.TUNE_SGM_MAX_STRIP, strip)
    for _ in range(2): pipe.run_device(left, right, hints)
    torch.cuda.synchronize()
    L.vppb200_stage_timing(1)
    for _ in range(4): out = pipe.run_device(left, right, hints)
    torch.cuda.synchronize()
    ms = (ctypes.c_float * 10)(); calls = ctypes.c_int(0)
    L.vppb200_stage_times(ms, ctypes.byref(calls)); L.vppb200_stage_timing(0)
    print(strip, "v_down %.3f v_up %.3f total_rsgm %.3f" % (ms[4] / calls.value, ms[5] / calls.value, sum(ms) / calls.value), float(out.mean()))
