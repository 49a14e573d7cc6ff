"""GPU timing of VPP maxDistance (BASELINE.json configs[2] shape and the K batch): row wavefront vs serial scan."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vppstereo_b200 import _lib, synth, vpp_core_opt as core

def run(shape, N, wave, reps):
    frames = [synth.make_pair(f, shape=shape, hints="random") for f in range(N)]
    l = torch.from_numpy(np.stack([p["left"] for p in frames])).cuda()
    r = torch.from_numpy(np.stack([p["right"] for p in frames])).cuda()
    g = torch.from_numpy(np.stack([p["hints"] for p in frames])).cuda()
    occ = torch.zeros(g.shape, dtype=torch.uint8, device="cuda")
    H, W = g.shape[-2:]
    _lib.set_tuning(_lib.TUNE_VPP_MD_WAVE, wave)
    ts = []
    for i in range(reps + 1):
        lt, rt = l.clone(), r.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        core.virtual_projection_scan_max_dist(lt, rt, g, W, H, 3, 0, 3, 64, 3, 1, 0.4, 0.0, occ, 0, 1, arith=1)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    _lib.set_tuning(_lib.TUNE_VPP_MD_WAVE, 1)
    hints = int((g > 0).sum())
    print(f"shape {shape} x{N} wave={wave}: {min(ts[1:]):.2f} ms (first {ts[0]:.2f}), {hints} hints, "
          f"{hints / min(ts[1:]) / 1e3:.2f} Mhints/s", flush=True)
    return lt, rt

if __name__ == "__main__":
    if "--k64" in sys.argv:
        run("K", 64, 1, 1)
        sys.exit(0)
    if "--scale" in sys.argv:
        for nb in (24, 48, 72):
            run("M", nb, 1, 1)
        sys.exit(0)
    if "--batch" in sys.argv:
        run("M", 1, 1, 2); run("M", 12, 1, 2); run("K", 64, 1, 3); run("V", 1, 1, 3)
        sys.exit(0)
    a = run("M", 1, 1, 3)
    if "--serial" in sys.argv:
        b = run("M", 1, 0, 1)
        print("wave == serial:", bool((a[0] == b[0]).all() and (a[1] == b[1]).all()))
    run("K", 64, 1, 3)
    run("V", 1, 1, 3)
