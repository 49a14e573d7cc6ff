"""CPU: host-side logic -- the C-ABI library loads and exports every symbol the header declares, the reference-shaped
front-ends validate arguments exactly like the reference wrapper (before touching any device), compute calls fail loudly
without a GPU (there is no CPU fallback), and frame sharding covers every frame exactly once."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_header_symbols():
    from vppstereo_b200 import _lib, build
    build.build()
    header = open(os.path.join(ROOT, "include", "vppstereo_b200.h")).read()
    declared = set(re.findall(r"\b(vppb200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared:
        assert hasattr(lib, s), f"{s} not exported"
    lib.vppb200_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.vppb200_version()


def test_tuning_keys_match_header():
    """every VPPB200_TUNE_* key of include/vppstereo_b200.h has the same value in the Python binding, and the library accepts it"""
    import re
    from vppstereo_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "vppstereo_b200.h")).read()
    keys = dict((m.group(1), int(m.group(2))) for m in re.finditer(r"#define\s+VPPB200_TUNE_(\w+)\s+(\d+)", hdr))
    assert len(keys) >= 11 and len(set(keys.values())) == len(keys)
    for name, value in keys.items():
        assert getattr(_lib, "TUNE_" + name) == value, name
    L = _lib.lib()
    assert L.vppb200_set_tuning(max(keys.values()) + 1, 0) != 0          # unknown key: VPPB200_ERR_ARG
    for name, value in keys.items():
        default = 0 if name in ("SGM_MAX_STRIP", "SGM_CLUSTERS", "SGM_BYTE_SUMS", "SGM_FUSE_COST", "RCP_HOST") else 1
        assert L.vppb200_set_tuning(value, default) == 0, name          # (host-side switches only: no GPU needed)


def test_no_torch_types_in_header():
    header = open(os.path.join(ROOT, "include", "vppstereo_b200.h")).read()
    assert "torch" not in header.lower() and "at::" not in header and "#include <cuda" not in header


def test_status_codes_without_gpu():
    """validation happens before any CUDA call, so it is testable here through the raw C ABI"""
    from vppstereo_b200 import _lib
    L = _lib.lib()
    one = ctypes.c_void_p(16)         # never dereferenced: validation fails first
    assert L.vppb200_census5x5(one, one, 24, 8, 1, None) == _lib.ERR_WIDTH
    assert L.vppb200_cost_census5x5_xyd(one, one, one, 32, 8, 12, 1, 1, None) == _lib.ERR_DISP
    assert L.vppb200_cost_census5x5_xyd(one, one, one, 32, 8, 264, 1, 1, None) == _lib.ERR_DISP
    assert L.vppb200_cost_census5x5_xyd(one, one, one, 32, 8, 8, 3, 1, None) == _lib.ERR_THREADS
    assert L.vppb200_match_wta(one, one, 32, 8, 8, ctypes.c_float(0.0), 1, None) == _lib.ERR_UNIQUENESS
    assert L.vppb200_match_wta_right(one, one, 32, 8, 8, ctypes.c_float(1.5), 1, None) == _lib.ERR_UNIQUENESS
    assert L.vppb200_subpixel_refine(one, one, 32, 8, 8, 2, None, 1, None) == _lib.ERR_METHOD
    assert L.vppb200_census5x5(None, one, 32, 8, 1, None) == _lib.ERR_ARG
    assert L.vppb200_compute_rsgm(one, one, one, None, None, one, 8, 32, 3, 12, 1, None, None, ctypes.c_size_t(0), 1, None) == _lib.ERR_DISP
    assert L.vppb200_compute_rsgm(one, one, one, None, None, one, 8, 32, 3, 16, 1, None, None, ctypes.c_size_t(0), 1, None) == _lib.ERR_WORKSPACE
    assert L.vppb200_rsgm_workspace_bytes(375, 1242, 3, 192, 1) > 2 * 1248 * 384 * 192
    assert L.vppb200_vpp_workspace_bytes(375, 1242, 3, 1) >= 2 * 375 * 1242


def test_front_end_errors_match_reference():
    from vppstereo_b200 import pyrSGM, rsgm, vpp_standalone
    z8, z32 = np.zeros((8, 24), np.uint8), np.zeros((8, 24), np.uint32)
    with pytest.raises(TypeError, match="multiple of 16"):
        pyrSGM.census5x5_SSE(z8, z32, 24, 8)
    with pytest.raises(TypeError, match="multiple of 16"):
        pyrSGM.median3x3_SSE(z8.astype(np.float32), z8.astype(np.float32), 24, 8)
    a32 = np.zeros((8, 32), np.uint32)
    with pytest.raises(TypeError, match="multiple of 8"):
        pyrSGM.costMeasureCensus5x5_xyd_SSE(a32, a32, np.zeros((8, 32, 12), np.uint16), 32, 8, 12, 1)
    with pytest.raises(TypeError, match="NumThreads"):
        pyrSGM.costMeasureCensus5x5_xyd_SSE(a32, a32, np.zeros((8, 32, 8), np.uint16), 32, 8, 8, 8)
    with pytest.raises(TypeError, match="Uniqueness"):
        pyrSGM.matchWTARight_SSE(np.zeros((8, 32, 8), np.uint16), np.zeros((8, 32), np.float32), 32, 8, 8, 0.0)
    with pytest.raises(TypeError, match="method"):
        pyrSGM.subPixelRefine(np.zeros((8, 32, 8), np.uint16), np.zeros((8, 32), np.float32), 32, 8, 8, 5)
    img = np.zeros((20, 30, 3), np.uint8)
    with pytest.raises(Exception, match="dmax % 8"):
        rsgm.compute_rsgm(img, img, img, dmax=100)                       # models/rsgm/rsgm.py:31-32
    with pytest.raises(Exception, match="dmax > 256"):
        rsgm.compute_rsgm(img, img, img, dmax=512)                       # models/rsgm/rsgm.py:34-35
    with pytest.raises(AssertionError):
        vpp_standalone.vpp(img, img, np.zeros((20, 30), np.float32), method="nope")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    from vppstereo_b200 import rsgm, vpp_standalone, pyrSGM
    img = np.zeros((20, 32, 3), np.uint8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        rsgm.compute_rsgm(img, img, img, dmax=16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        vpp_standalone.vpp(img, img, np.ones((20, 32), np.float32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pyrSGM.census5x5_SSE(np.zeros((8, 32), np.uint8), np.zeros((8, 32), np.uint32), 32, 8)


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under vppstereo_b200/ may reference it"""
    pkg = os.path.join(ROOT, "vppstereo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} mentions the oracle"


def test_shard_ranges_cover_all_frames():
    from vppstereo_b200 import dist
    for n in (1, 7, 64, 1024, 1000):
        for world in (1, 2, 3, 4, 8):
            if n < world:
                continue
            ranges = [dist.shard_range(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1 and sizes == dist.shard_sizes(n, world)
    assert dist.chunks(3, 20, 8) == [(3, 11), (11, 19), (19, 20)]


def test_glibc_generator_state_continues():
    """consecutive scans keep consuming one stream, like the process-global libc state (vpp_core_opt.pyx:33-35)"""
    from vppstereo_b200 import vpp_core_opt as core
    core.init_rand(42); a = core.draw_pattern(1000)
    core.init_rand(42); b = np.concatenate([core.draw_pattern(300), core.draw_pattern(700)])
    assert np.array_equal(a, b)
    libc = ctypes.CDLL("libc.so.6")
    libc.srand(42)
    assert np.array_equal(a[:64], np.array([libc.rand() % 256 for _ in range(64)], np.uint8))


def test_patch_threshold_tabulation():
    """vpp(use_distance_patch=True): the device evaluates the reference's patch-size function (float32 ratio, libm pow, round
    half to even; vpp_standalone.py:6-9) through per-frame disparity thresholds found on the host; the thresholds must
    reproduce the direct evaluation for every disparity, including the neighbours of each step."""
    from vppstereo_b200 import vpp_standalone as vs
    rng = np.random.default_rng(0)
    for dmin, dmax, wsize, gamma in [(1.5, 190.25, 7, 0.3), (0.5, 3.0, 9, 0.3), (10.0, 10.5, 5, 1.0), (2.0, 64.0, 3, 0.1), (1.0, 200.0, 1, 0.3)]:
        dmin, dmax = np.float32(dmin), np.float32(dmax)
        thr = vs._patch_thresholds(dmin, dmax, wsize, gamma)
        assert thr.shape == (wsize - 1,) and (np.diff(thr) >= 0).all()
        ds = np.concatenate([rng.uniform(dmin, dmax, 2000).astype(np.float32), [dmin, dmax]]).astype(np.float32)
        for t in thr[np.isfinite(thr)]:
            ds = np.concatenate([ds, [np.nextafter(t, np.float32(0)), t, np.nextafter(t, np.float32(1e9))]]).astype(np.float32)
        ds = ds[(ds >= dmin) & (ds <= dmax)]
        for d in ds:
            assert 1 + int((d >= thr).sum()) == vs._patch_size(d, dmin, dmax, wsize, gamma), (d, dmin, dmax, wsize, gamma)
    with pytest.raises(ZeroDivisionError):
        vs._patch_thresholds(np.float32(4.0), np.float32(4.0), 5, 0.3)


def test_band_with_halo():
    from vppstereo_b200 import dist as vd
    for H, world, halo in [(2000, 8, 3), (375, 4, 1), (10, 3, 5)]:
        covered = []
        for r in range(world):
            (lo, hi), (rlo, rhi) = vd.band_with_halo(H, r, world, halo)
            assert 0 <= rlo <= lo <= hi <= rhi <= H and lo - rlo <= halo and rhi - hi <= halo
            assert rlo == max(lo - halo, 0) and rhi == min(hi + halo, H)
            covered += list(range(lo, hi))
        assert covered == list(range(H))


def test_default_rcp_table_is_the_recorded_intel_table(orc, golden_lut):
    """subPixelRefine's reciprocal (RSGM/StereoBMHelper.cpp:752-756): the library's default is the FIXED Intel table (2048
    mantissas, csrc/rcp_intel_table.inc) = the table recorded with the golden vectors; VPPB200_TUNE_RCP_HOST selects the host
    CPU's own RCPSS, which is what the oracle (like the reference) evaluates."""
    from vppstereo_b200 import _lib
    L = _lib.lib()
    lut = np.empty(65536, np.float32)
    try:
        _lib.set_tuning(_lib.TUNE_RCP_HOST, 0)
        assert L.vppb200_rcp_lut_host(lut.ctypes.data_as(ctypes.c_void_p)) == 0
        assert np.array_equal(lut.view(np.uint32), golden_lut.view(np.uint32))
        _lib.set_tuning(_lib.TUNE_RCP_HOST, 1)
        assert L.vppb200_rcp_lut_host(lut.ctypes.data_as(ctypes.c_void_p)) == 0
        assert np.array_equal(lut.view(np.uint32), orc.rcp_lut().view(np.uint32))
    finally:
        _lib.set_tuning(_lib.TUNE_RCP_HOST, 0)


def test_device_pattern_is_a_pure_counter_hash():
    """vpp_core_opt.device_pattern restates csrc/vpp.cu::counter_pattern: prefix-stable, frame- and seed-dependent, uniform bytes"""
    from vppstereo_b200 import vpp_core_opt as core
    a = core.device_pattern(12345, 3, 4096)
    assert a.dtype == np.uint8 and np.array_equal(a[:100], core.device_pattern(12345, 3, 100))
    assert not np.array_equal(a, core.device_pattern(12345, 4, 4096)) and not np.array_equal(a, core.device_pattern(12346, 3, 4096))
    assert 100 < a.mean() < 155 and len(np.unique(a)) == 256
    # scalar restatement of the same hash
    def one(seed, f, i):
        k = (seed ^ (f * 0x9E3779B97F4A7C15)) & (2**64 - 1)
        key = (k & 0xFFFFFFFF) ^ (((k >> 32) * 0x85EBCA6B) & 0xFFFFFFFF)
        h = ((i * 0x9E3779B1) & 0xFFFFFFFF) ^ key
        h ^= h >> 16; h = (h * 0x85EBCA6B) & 0xFFFFFFFF
        h ^= h >> 13; h = (h * 0xC2B2AE35) & 0xFFFFFFFF
        h ^= h >> 16
        return h >> 24
    assert [one(2**63 + 5, 63, i) for i in (0, 1, 77, 4095)] == [int(core.device_pattern(2**63 + 5, 63, 4096)[i]) for i in (0, 1, 77, 4095)]


def test_pipeline_rejects_bad_operands_without_gpu():
    """VppRsgmPipeline._check validates before any pointer reaches a kernel (exercised here on a stand-in object: no GPU needed)"""
    import torch
    from vppstereo_b200.pipeline import VppRsgmPipeline
    p = VppRsgmPipeline.__new__(VppRsgmPipeline)
    p.torch, p.H, p.W, p.C, p.N, p.device = torch, 8, 16, 3, 4, torch.device("cuda", 0)
    p.vpp_stream = p.main_stream = p.tail_stream = None; p._stream_sets = None
    l = torch.zeros((2, 8, 16, 3), dtype=torch.uint8); g = torch.zeros((2, 8, 16), dtype=torch.float32)
    assert p._check(l, l, g, host=True) == 2
    with pytest.raises(ValueError, match="CUDA tensor"):
        p._check(l, l, g)                                               # host tensors on the device path
    with pytest.raises(TypeError, match="float32"):
        p._check(l, l, g.double(), host=True)
    with pytest.raises(ValueError, match="shape"):
        p._check(l, l[:1], g, host=True)
    with pytest.raises(ValueError, match="contiguous"):
        p._check(l, l, torch.zeros((2, 16, 8), dtype=torch.float32).transpose(1, 2), host=True)
    with pytest.raises(ValueError, match="does not fit"):
        p._check(torch.zeros((5, 8, 16, 3), dtype=torch.uint8), l, g, host=True)
    with pytest.raises(TypeError, match="torch tensor"):
        p._check(l.numpy(), l, g, host=True)
