"""Drop-in for vpp_standalone.py: `vpp(left, right, gt, ...) -> (lc, rc)` with the reference's signature and numba
arithmetic (everything float64, round half to even; vpp_standalone.py:14-369), running on the device.

    from vppstereo_b200.vpp_standalone import vpp        # instead of  from vpp_standalone import vpp   (test.py:16)

numpy in -> numpy copies out (inputs untouched, vpp_standalone.py:397); CUDA tensors in -> CUDA tensors out.  The
reference's numba generator is never seeded by test.py, so its random pattern is not reproducible;
here the pattern comes from `pattern=` (extra keyword, uint8 stream in draw order) or from a torch generator.
`vpp_to_network` adds the device hand-off of test.py:179-197 (uint8 HWC -> float32 CHW / 255, replicate pad).
"""
import numpy as np

from . import _lib
from . import vpp_core_opt as _core

__all__ = ["vpp", "vpp_to_network"]


def vpp(left, right, gt, wsize=3, wsizeAgg_x=64, wsizeAgg_y=3, left2right=True, blending=0.4, use_distance_patch=False,
        use_bilateral_patch=False, distance_gamma=0.3, bilateral_o_xy=2, bilateral_o_i=1, bilateral_th=.001,
        uniform_color=False, method="rnd", c_occ=0.00, g_occ=None, discard_occ=False, interpolate=True, pattern=None,
        seed=None):
    assert method in ["rnd", "maxDistance"]                                  # vpp_standalone.py:400
    if use_distance_patch or use_bilateral_patch:
        raise NotImplementedError("distance / bilateral adaptive patches (vpp_standalone.py:6-11,:371-394) are the "
                                  "next scope row (SURVEY.md 8f-2) and are not built yet")
    torch = _lib.require_cuda()
    host = not _lib.is_tensor(left)
    lc = _lib.as_device(left, torch.uint8).clone()
    rc = _lib.as_device(right, torch.uint8).clone()
    g = _lib.as_device(gt, torch.float32)                                    # gt.astype(np.float32) (:398)
    direction = 1 if left2right else 0
    batched = g.dim() == 3
    if lc.dim() == g.dim():                                                  # gray: add the channel axis (:403-404)
        lc, rc = lc.unsqueeze(-1), rc.unsqueeze(-1)
    lc, rc = lc.contiguous(), rc.contiguous()
    H, W = g.shape[-2:]
    Cn = lc.shape[-1]
    if True:   # frames without points are left untouched by the scan itself (vpp_standalone.py:407-408): no host sync needed
        if g_occ is None:
            occ = torch.zeros(g.shape, dtype=torch.uint8, device=g.device)   # np.zeros_like(gt) (:424-425)
        else:
            occ = (_lib.as_device(g_occ, torch.float32) != 0).to(torch.uint8)
        if method == "maxDistance":
            _core._scan("max_dist", lc, rc, g, W, H, Cn, uniform_color, wsize, (wsizeAgg_x, wsizeAgg_y), direction,
                        blending, c_occ, occ, discard_occ, interpolate, None, 1, want_counts=False)
        else:
            # no pattern given: counter-based generation inside the splat kernel (the reference's numba RNG is unseeded)
            rng_seed = None
            if pattern is None:
                rng_seed = int(seed) if seed is not None else int.from_bytes(__import__("os").urandom(8), "little")
                rng_seed |= 1 << 63                  # a given seed of 0 must still select the device generator
            _core._scan("rnd", lc, rc, g, W, H, Cn, uniform_color, wsize, None, direction, blending, c_occ, occ,
                        discard_occ, interpolate, pattern, 1, device_rng_seed=rng_seed, want_counts=False)
    if host:
        return lc.cpu().numpy(), rc.cpu().numpy()
    return lc, rc


def vpp_to_network(img_u8, pad_to=32):
    """uint8 [H,W,C] / [N,H,W,C] CUDA tensor -> float32 [N,C,H+pad,W+pad] = float32(u8/255.) with the symmetric
    replicate padding of test.py:187-197 (pad_to=None: no padding).  Returns (tensor, (left,right,top,bottom))."""
    torch = _lib.require_cuda()
    t = _lib.as_device(img_u8, torch.uint8)
    if t.dim() == 3:
        t = t[None]
    t = t.contiguous()
    N, H, W, Cn = t.shape
    pad_h = pad_w = 0
    if pad_to:
        pad_h = (((H // pad_to) + 1) * pad_to - H) % pad_to
        pad_w = (((W // pad_to) + 1) * pad_to - W) % pad_to
    pl, pr, pt, pb = pad_w // 2, pad_w - pad_w // 2, pad_h // 2, pad_h - pad_h // 2
    out = torch.empty((N, Cn, H + pad_h, W + pad_w), dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        rc = _lib.lib().vppb200_u8hwc_to_f32chw(_lib.ptr(t), _lib.ptr(out), H, W, Cn, pt, pb, pl, pr, N,
                                                 _lib.stream_ptr(t.device))
    _lib.check(rc, "vpp_to_network")
    return out, (pl, pr, pt, pb)
