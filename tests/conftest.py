import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _host_rcpss_for_oracle_parity():
    """The oracle follows the reference and evaluates the sub-pixel reciprocal with the HOST CPU's RCPSS instruction; the library
    defaults to the fixed Intel table (same numbers on every host).  The GPU parity tests compare against the oracle on whatever
    host the box has, so they switch the library to the host's RCPSS; tests of the default table switch back themselves."""
    try:
        import torch
        if torch.cuda.is_available():
            from vppstereo_b200 import _lib
            _lib.set_tuning(_lib.TUNE_RCP_HOST, 1)
    except Exception:
        pass
    yield


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (oracle/*.c through ctypes).  Test infrastructure only."""
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def golden_lut():
    return np.load(os.path.join(GOLDEN, "rcp_lut.npz"))["lut"]


def golden_rsgm_names():
    return ["tsukuba_crop", "kitti_crop", "kitti_crop_d72", "synth_colour", "synth_guided"]


def load_golden_rsgm(name):
    return dict(np.load(os.path.join(GOLDEN, f"rsgm_{name}.npz")))


@pytest.fixture(scope="session")
def golden_vpp():
    return dict(np.load(os.path.join(GOLDEN, "vpp_cases.npz")))


def bits(a):
    """float32 arrays are compared on their bit patterns (NaN-safe, -0.0 aware)."""
    a = np.ascontiguousarray(a)
    return a.view(np.uint32) if a.dtype == np.float32 else a


def assert_same(a, b, what):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    ne = bits(a) != bits(b)
    n = int(ne.sum())
    if n:
        idx = np.argwhere(ne)[:8].tolist()
        raise AssertionError(f"{what}: {n} of {a.size} elements differ; first at {idx}: "
                             f"{[a[tuple(i)].item() for i in idx[:4]]} vs {[b[tuple(i)].item() for i in idx[:4]]}")
