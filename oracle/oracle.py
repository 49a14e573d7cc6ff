"""ctypes front-end of the C oracle (oracle/rsgm_oracle.c, oracle/vpp_oracle.c, oracle/filter_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.
The product package (vppstereo_b200/) must never import this module.

`build()` compiles oracle/_build/liboracle.so with gcc (no FMA contraction: the reference's VPP build has none).
Function names mirror the reference operators they restate (file:line in the C sources).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(BUILD_DIR, "liboracle.so")
SOURCES = [os.path.join(HERE, f) for f in ("rsgm_oracle.c", "vpp_oracle.c", "filter_oracle.c")]
_lib = None


def build(force=False):
    os.makedirs(BUILD_DIR, exist_ok=True)
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in SOURCES):
        return LIB_PATH
    cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
    cmd = [cc, "-O2", "-fPIC", "-shared", "-std=gnu11", "-ffp-contract=off", "-msse2", "-fvisibility=hidden",
           "-o", LIB_PATH] + SOURCES + ["-lm"]
    subprocess.check_call(cmd)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.orc_rcp_nz.restype = C.c_float
        _lib.orc_rcp_nz.argtypes = [C.c_float]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


# ---------------------------------------------------------------- third-party restatements
def rgb2gray(img):
    img = _c(img, np.uint8)
    out = np.empty(img.shape[:2], np.uint8)
    lib().orc_rgb2gray(_p(img), _p(out), C.c_int(out.size))
    return out


def bgr2gray(img):
    img = _c(img, np.uint8)
    out = np.empty(img.shape[:2], np.uint8)
    lib().orc_bgr2gray(_p(img), _p(out), C.c_int(out.size))
    return out


def pad_reflect(img, top, bottom, left, right):
    img = np.ascontiguousarray(img)
    H, W = img.shape[:2]
    elem = img.itemsize * (img.shape[2] if img.ndim == 3 else 1)
    out = np.empty((H + top + bottom, W + left + right) + img.shape[2:], img.dtype)
    lib().orc_pad_reflect(_p(img), _p(out), H, W, elem, top, bottom, left, right)
    return out


def filter_speckles_u8(img, new_val=0, max_size=200, max_diff=10):
    img = _c(img, np.uint8).copy()
    lib().orc_filter_speckles_u8(_p(img), img.shape[0], img.shape[1], new_val, max_size, max_diff)
    return img


# ---------------------------------------------------------------- pyrSGM operators (same names / argument order)
def census5x5_SSE(src, dst, W, H):
    src = _c(src, np.uint8)
    assert dst.dtype == np.uint32 and dst.flags.c_contiguous
    if W % 16:
        raise TypeError("Width must be a multiple of 16")
    lib().orc_census5x5(_p(src), _p(dst), int(W), int(H))


def costMeasureCensus5x5_xyd_SSE(cl, cr, dsi, W, H, D, nthreads=1):
    if W % 16 or D % 8 or D > 256 or nthreads not in (1, 2, 4):
        raise TypeError("bad arguments")
    cl, cr = _c(cl, np.uint32), _c(cr, np.uint32)
    assert dsi.dtype == np.uint16 and dsi.flags.c_contiguous
    lib().orc_cost_census(_p(cl), _p(cr), _p(dsi), int(W), int(H), int(D))


def aggregate_SSE(img, dsi, dsiAgg, W, H, D, P1=7, P2min=17, Alpha=0.25, Gamma=50, honor_params=False):
    """The reference ignores P1..Gamma (RSGM/pyrSGM.cpp:519 vs :557-560); so does this unless honor_params."""
    if W % 16 or D % 8 or D > 256:
        raise TypeError("bad arguments")
    img = np.ascontiguousarray(img, np.uint8).reshape(-1)[: W * H].copy()
    dsi = _c(dsi, np.uint16)
    assert dsiAgg.dtype == np.uint16 and dsiAgg.flags.c_contiguous
    p = (int(P1), int(P2min), float(Alpha), int(Gamma)) if honor_params else (7, 17, 0.25, 50)
    lib().orc_sgm_aggregate(_p(img), _p(dsi), _p(dsiAgg), int(W), int(H), int(D),
                            C.c_int(p[0]), C.c_int(p[1]), C.c_float(p[2]), C.c_int(p[3]))


def matchWTA_SSE(dsiAgg, disp, W, H, D, uniqueness=0.95):
    if W % 16 or D % 8 or D > 256 or not (0.0 < uniqueness <= 1.0):
        raise TypeError("bad arguments")
    lib().orc_wta_left(_p(_c(dsiAgg, np.uint16)), _p(disp), int(W), int(H), int(D))


def matchWTARight_SSE(dsiAgg, disp, W, H, D, uniqueness=0.95):
    if W % 16 or D % 8 or D > 256 or not (0.0 < uniqueness <= 1.0):
        raise TypeError("bad arguments")
    lib().orc_wta_right(_p(_c(dsiAgg, np.uint16)), _p(disp), int(W), int(H), int(D))


def rcp_lut():
    """65536-entry table of rcp_nz_ss(-2k) on THIS host CPU (RSGM/StereoBMHelper.cpp:752-756)."""
    lut = np.empty(65536, np.float32)
    lib().orc_rcp_lut(_p(lut))
    return lut


def subPixelRefine(dsi, disp, W, H, D, method=0, lut=None):
    if W % 16 or D % 8 or D > 256 or method not in (0, 1):
        raise TypeError("bad arguments")
    assert disp.dtype == np.float32 and disp.flags.c_contiguous
    lp = _p(_c(lut, np.float32)) if lut is not None else None
    lib().orc_subpixel(_p(_c(dsi, np.uint16)), _p(disp), int(W), int(H), int(D), int(method), lp)


def median3x3_SSE(src, dst, W, H):
    if W % 16:
        raise TypeError("Width must be a multiple of 16")
    assert dst.dtype == np.float32 and dst.flags.c_contiguous
    lib().orc_median3x3(_p(_c(src, np.float32)), _p(dst), int(W), int(H))


# ---------------------------------------------------------------- rsgm.py tail
def linear_interpolate(dmap, n=15, th=3.0):
    assert dmap.dtype == np.float32 and dmap.flags.c_contiguous
    lib().orc_linear_interpolate(_p(dmap), dmap.shape[1], dmap.shape[0], int(n), C.c_float(th))


def left_right_check(dl, dr, th=1.0):
    dl, dr = _c(dl, np.float32), _c(dr, np.float32)
    mask = np.empty(dl.shape, np.uint8)
    lib().orc_lr_check(_p(dl), _p(dr), _p(mask), dl.shape[1], dl.shape[0], C.c_float(th))
    return mask


def interpolate_background(dmap):
    assert dmap.dtype == np.float32 and dmap.flags.c_contiguous
    lib().orc_interpolate_background(_p(dmap), dmap.shape[1], dmap.shape[0])


def guided_dsi(dsi, hints, validhints):
    dsi = _c(dsi, np.uint16).copy()
    H, W, D = dsi.shape
    lib().orc_guided_dsi(_p(dsi), _p(_c(hints, np.float32)), _p(_c(validhints, np.float32)), W, H, D)
    return dsi


def compute_rsgm(left, left_vpp, right_vpp, hints=None, validhints=None, dmax=192, p1=11, p2min=17, alpha=0.5,
                 gamma=35, uniqueness=0.95, subpixel=True, rcp_lut_override=None):
    """models/rsgm/rsgm.py:250-294 (p1..gamma are accepted and ignored, as in the reference)."""
    if dmax % 8 or dmax > 256:
        raise Exception(f"Invalid dmax ({dmax})")
    left, left_vpp, right_vpp = (_c(a, np.uint8) for a in (left, left_vpp, right_vpp))
    H, W = left.shape[:2]
    Cn = left_vpp.shape[2] if left_vpp.ndim == 3 else 1
    if left.ndim == 3 and left.shape[2] != Cn or left.ndim == 2 and Cn != 1:
        raise ValueError("left / left_vpp channel mismatch is not supported by the oracle")
    out = np.empty((H, W), np.float32)
    hp = _p(_c(hints, np.float32)) if hints is not None and validhints is not None else None
    vp = _p(_c(validhints, np.float32)) if hints is not None and validhints is not None else None
    lp = _p(_c(rcp_lut_override, np.float32)) if rcp_lut_override is not None else None
    rc = lib().orc_compute_rsgm(_p(left), _p(left_vpp), _p(right_vpp), H, W, Cn, int(dmax), int(bool(subpixel)),
                                lp, hp, vp, _p(out))
    if rc:
        raise Exception(f"oracle compute_rsgm failed ({rc})")
    return out


# ---------------------------------------------------------------- vpp_core_opt operators
def libc_rand_stream(seed, n):
    """`init_rand(seed)` then n draws of `rand() % 256` (vpp_core_opt.pyx:33-35,:93,:102): glibc's process-global generator."""
    libc = C.CDLL("libc.so.6")
    libc.srand(C.c_uint(int(seed) & 0xFFFFFFFF))
    out = np.empty(n, np.uint8)
    rand = libc.rand
    for i in range(n):
        out[i] = rand() % 256
    return out


def stream_length(g, wsize, channels, uniform_color):
    """Number of pattern draws one scan consumes."""
    g = np.asarray(g)
    H, W = g.shape
    n = (wsize - 1) // 2
    ys, xs = np.nonzero(g > 0)
    if uniform_color:
        return int(len(ys)) * channels
    ny = np.minimum(ys + n, H - 1) - np.maximum(ys - n, 0) + 1
    nx = np.minimum(xs + n, W - 1) - np.maximum(xs - n, 0) + 1
    return int((ny * nx).sum()) * channels


def virtual_projection_scan_rnd(l, r, g, width, height, channels, uniform_color, wsize, direction, c, c_occ, g_occ,
                                discard_occluded, interpolate, stream=None, mode=0):
    assert l.dtype == np.uint8 and r.dtype == np.uint8 and l.flags.c_contiguous and r.flags.c_contiguous
    g = _c(g, np.float32)
    g_occ = _c(g_occ, np.uint8)
    stream = _c(stream if stream is not None else np.zeros(0, np.uint8), np.uint8)
    used = C.c_long(0)
    n = lib().orc_vpp_scan_rnd(_p(l), _p(r), _p(g), int(width), int(height), int(channels), int(bool(uniform_color)),
                               int(wsize), int(direction), C.c_double(c), C.c_double(c_occ), _p(g_occ),
                               int(bool(discard_occluded)), int(bool(interpolate)), int(mode), _p(stream),
                               C.c_long(stream.size), C.byref(used))
    return n


def virtual_projection_scan_max_dist(l, r, g, width, height, channels, uniform_color, wsize, wsize_agg_x, wsize_agg_y,
                                     direction, c, c_occ, g_occ, discard_occluded, interpolate, mode=0):
    assert l.dtype == np.uint8 and r.dtype == np.uint8 and l.flags.c_contiguous and r.flags.c_contiguous
    g = _c(g, np.float32)
    g_occ = _c(g_occ, np.uint8)
    return lib().orc_vpp_scan_max_dist(_p(l), _p(r), _p(g), int(width), int(height), int(channels),
                                       int(bool(uniform_color)), int(wsize), int(wsize_agg_x), int(wsize_agg_y),
                                       int(direction), C.c_double(c), C.c_double(c_occ), _p(g_occ),
                                       int(bool(discard_occluded)), int(bool(interpolate)), int(mode))


def gt_reshape(gt):
    gt = _c(gt, np.float32)
    H, W = gt.shape
    out = np.zeros((W * H, 4), np.float32)
    n = lib().orc_gt_reshape(_p(gt), W, H, _p(out))
    return out[:n]


def bilateral_filling(dmap, img, n, o_xy=2, o_i=1, th=.001):
    """vpp_standalone.py:371-394."""
    dmap = _c(dmap, np.float32)
    img = _c(img, np.uint8)
    H, W = dmap.shape
    assert img.shape == dmap.shape
    out = np.empty((H, W), np.float32)
    lib().orc_bilateral_filling(_p(dmap), _p(img), W, H, int(n), C.c_double(o_xy), C.c_double(o_i), C.c_double(th), _p(out))
    return out


def vpp(left, right, gt, wsize=3, wsizeAgg_x=64, wsizeAgg_y=3, left2right=True, blending=0.4, use_distance_patch=False,
        use_bilateral_patch=False, distance_gamma=0.3, bilateral_o_xy=2, bilateral_o_i=1, bilateral_th=.001, uniform_color=False,
        method="rnd", c_occ=0.0, g_occ=None, discard_occ=False, interpolate=True, stream=None, mode=1):
    """vpp_standalone.py:396-432 (mode 1 = numba arithmetic as test.py uses it; the adaptive-patch flags exist in numba only)."""
    lc, rc = np.copy(left), np.copy(right)
    gt = gt.astype(np.float32)
    assert method in ["rnd", "maxDistance"]
    direction = 1 if left2right else 0
    if lc.ndim < 3:
        lc, rc = np.expand_dims(lc, -1), np.expand_dims(rc, -1)
    if np.count_nonzero(gt) == 0:
        return lc, rc
    lc, rc = np.ascontiguousarray(lc), np.ascontiguousarray(rc)
    if g_occ is None:
        g_occ = np.zeros(gt.shape, np.uint8)
    g_occ = (np.asarray(g_occ) != 0).astype(np.uint8)
    H, W, Cn = lc.shape
    if use_distance_patch or use_bilateral_patch:
        assert mode == 1
        dmin, dmax = gt[gt > 0].min(), gt[gt > 0].max()
        if use_distance_patch and dmin == dmax:
            raise ZeroDivisionError("division by zero")               # numba's python error model (:7)
        gray = bgr2gray(lc) if Cn == 3 else np.ascontiguousarray(lc[..., 0])
        filled = bilateral_filling(gt, gray, (wsize - 1) // 2, bilateral_o_xy, bilateral_o_i, bilateral_th) if use_bilateral_patch \
            else gt.copy()
        gt_c, filled, occ_c = _c(gt, np.float32), _c(filled, np.float32), _c(g_occ, np.uint8)
        if method == "maxDistance":
            lib().orc_vpp_scan_max_dist_adaptive(_p(lc), _p(rc), _p(gt_c), _p(filled), W, H, Cn, int(bool(uniform_color)), int(wsize),
                                                 int(wsizeAgg_x), int(wsizeAgg_y), direction, C.c_double(blending), C.c_double(c_occ),
                                                 _p(occ_c), int(bool(discard_occ)), int(bool(interpolate)), int(bool(use_distance_patch)),
                                                 int(bool(use_bilateral_patch)), C.c_float(dmin), C.c_float(dmax), C.c_double(distance_gamma))
        else:
            st = _c(stream if stream is not None else np.zeros(0, np.uint8), np.uint8)
            used = C.c_long(0)
            lib().orc_vpp_scan_rnd_adaptive(_p(lc), _p(rc), _p(gt_c), _p(filled), W, H, Cn, int(bool(uniform_color)), int(wsize),
                                            direction, C.c_double(blending), C.c_double(c_occ), _p(occ_c), int(bool(discard_occ)),
                                            int(bool(interpolate)), _p(st), C.c_long(st.size), C.byref(used),
                                            int(bool(use_distance_patch)), int(bool(use_bilateral_patch)), C.c_float(dmin),
                                            C.c_float(dmax), C.c_double(distance_gamma))
        return lc, rc
    if method == "maxDistance":
        virtual_projection_scan_max_dist(lc, rc, gt, W, H, Cn, uniform_color, wsize, wsizeAgg_x, wsizeAgg_y, direction,
                                         blending, c_occ, g_occ, discard_occ, interpolate, mode=mode)
    else:
        virtual_projection_scan_rnd(lc, rc, gt, W, H, Cn, uniform_color, wsize, direction, blending, c_occ, g_occ,
                                    discard_occ, interpolate, stream=stream, mode=mode)
    return lc, rc


# ---------------------------------------------------------------- filter.py
def occlusion_heuristic(dmap, rx=9, ry=7, l=2, g=0.4375, th_conf=1, th_filter=0.1):
    """filter.py:246-292: (filtered + interpolated disparity map, binary occlusion mask: 0 = visible hint)."""
    dmap = _c(dmap, np.float32)
    H, W = dmap.shape
    out = np.empty((H, W), np.float32)
    conf = np.empty((H, W), np.uint8)
    lib().orc_occlusion_heuristic(_p(dmap), _p(out), _p(conf), W, H, int(rx), int(ry), C.c_double(l), C.c_double(g),
                                  C.c_double(th_conf), C.c_double(th_filter))
    return out, conf
