"""vppstereo_b200 -- B200-native (sm_100a) implementation of vppstereo's VPP + rSGM hot path.

Drop-in mirrors of the reference interfaces (same names, argument order and error behaviour):
  vppstereo_b200.pyrSGM          <-> thirdparty/stereo-vision/reconstruction/base/rSGM/pyrSGM.cpp:761-774
  vppstereo_b200.vpp_core_opt    <-> vpp_core/vpp_core_opt.pyx
  vppstereo_b200.rsgm            <-> models/rsgm/rsgm.py  (compute_rsgm)
  vppstereo_b200.vpp_standalone  <-> vpp_standalone.py    (vpp)
All compute runs in hand-written CUDA kernels behind the C-ABI of include/vppstereo_b200.h; there is no CPU path.
"""
__version__ = "0.1.0"
