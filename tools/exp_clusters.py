import sys, torch, numpy as np
sys.path.insert(0, '.')
from vppstereo_b200 import _lib, synth
from vppstereo_b200.pipeline import VppRsgmPipeline
dev = torch.device('cuda', 0)
B=64
frames=[synth.make_pair(f, shape="K", hints="lidar") for f in range(4)]
idx=[i%4 for i in range(B)]
left=torch.from_numpy(np.stack([frames[i]["left"] for i in idx])).to(dev)
right=torch.from_numpy(np.stack([frames[i]["right"] for i in idx])).to(dev)
hints=torch.from_numpy(np.stack([frames[i]["hints"] for i in idx])).to(dev)
pipe=VppRsgmPipeline(375,1242,3,batch=B,dmax=192,device=dev)
print(torch.cuda.get_device_properties(0).multi_processor_count)
for nc in (0, 16, 17, 18, 14, 8):
    _lib.set_tuning(_lib.TUNE_SGM_CLUSTERS, nc)
    for _ in range(2): pipe.run_device(left,right,hints)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4): out=pipe.run_device(left,right,hints)
    e1.record(); torch.cuda.synchronize()
    print(nc, e0.elapsed_time(e1)/4, float(out.mean()))
