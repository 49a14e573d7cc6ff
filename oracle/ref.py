"""Loader for the compiled reference under oracle/_ref/ (see build_ref.py).

TEST INFRASTRUCTURE ONLY: used by tests/, tests/golden/make_golden.py and bench.py's CPU-baseline legs.
The product package never imports this module.

`load()` returns a namespace with the reference's own callables:
  pyrSGM, vpp_core_opt           compiled extension modules (reference sources, unmodified)
  rsgm, vpp_standalone, filter   byte-compiled reference Python (models/rsgm/rsgm.py, vpp_standalone.py, filter.py)
The pyrSGM census reads uninitialised memory; callers that need determinism must run the
process with MALLOC_MMAP_THRESHOLD_=65536 set *before interpreter start* (tests spawn a subprocess for this).
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_cache = None


def available():
    return os.path.isdir(REF_DIR) and any(f.startswith("pyrSGM") and f.endswith(".so") for f in os.listdir(REF_DIR))


def _load_pyc(name, fname):
    path = os.path.join(REF_DIR, fname)
    loader = importlib.machinery.SourcelessFileLoader(name, path)
    spec = importlib.util.spec_from_loader(name, loader)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    loader.exec_module(mod)
    return mod


def load():
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError("oracle/_ref not built: run `python oracle/build_ref.py` where /root/reference exists")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import cv2
    cv2.setNumThreads(0)  # as dataloaders/frame_utils.py:7 does
    ns = types.SimpleNamespace()
    ns.pyrSGM = importlib.import_module("pyrSGM")          # rsgm_ref.pycode does `from pyrSGM import ...`
    ns.vpp_core_opt = importlib.import_module("vpp_core_opt")
    ns.rsgm = _load_pyc("rsgm_ref", "rsgm_ref.pycode")
    ns.vpp_standalone = _load_pyc("vpp_standalone_ref", "vpp_standalone_ref.pycode")
    ns.filter = _load_pyc("filter_ref", "filter_ref.pycode")
    _cache = ns
    return ns


class _RefModelsFinder:
    """meta-path finder: refmodels[.pkg[.module]] -> oracle/_ref/refmodels/.../*.pycode (sourceless, relative imports work)"""

    @staticmethod
    def find_spec(fullname, path=None, target=None):
        if fullname != "refmodels" and not fullname.startswith("refmodels."):
            return None
        base = os.path.join(REF_DIR, *fullname.split("."))
        if os.path.isdir(base):
            init = os.path.join(base, "__init__.pycode")
            loader = importlib.machinery.SourcelessFileLoader(fullname, init)
            return importlib.util.spec_from_file_location(fullname, init, loader=loader, submodule_search_locations=[base])
        if os.path.exists(base + ".pycode"):
            loader = importlib.machinery.SourcelessFileLoader(fullname, base + ".pycode")
            return importlib.util.spec_from_file_location(fullname, base + ".pycode", loader=loader)
        return None


def load_nets():
    """(RAFTStereo, PSMNet): the reference's own network classes (models/raft_stereo/raft_stereo.py:23, models/psmnet/psmnet.py:86)
    from the byte-compiled packages under oracle/_ref/refmodels.  They are the consumers of the projected images (BASELINE
    configs[3]); tests and bench.py instantiate them with random weights.  `opt_einsum` (imported, never called, by
    models/raft_stereo/update.py:4) is absent from this image: a stand-in module is registered."""
    if not os.path.isdir(os.path.join(REF_DIR, "refmodels")):
        raise RuntimeError("oracle/_ref/refmodels not built: run `python oracle/build_ref.py` where /root/reference exists")
    if not any(f is _RefModelsFinder for f in sys.meta_path):
        sys.meta_path.insert(0, _RefModelsFinder)
    if "opt_einsum" not in sys.modules:
        try:
            import opt_einsum  # noqa: F401
        except ImportError:
            import torch
            stub = types.ModuleType("opt_einsum")
            stub.contract = torch.einsum
            sys.modules["opt_einsum"] = stub
    raft = importlib.import_module("refmodels.raft_stereo.raft_stereo")
    psm = importlib.import_module("refmodels.psmnet.psmnet")
    return raft.RAFTStereo, psm.PSMNet


def load_losses():
    """the reference's losses.py (sample_hints, losses.py:5-10), byte-compiled"""
    return _load_pyc("losses_ref", "losses_ref.pycode")


def zero_unwritten_census(ct):
    """The reference never writes census rows 0,1,H-2,H-1 nor (H-3, {W-16,W-15,W-2,W-1}) (RSGM/FastFilters.cpp:181-442)
    and returns whatever malloc held there (non-deterministic, feeds the cost volume through row H-3).  Parity is
    defined with those pixels = 0; this helper pins the reference's output to that definition."""
    H, W = ct.shape
    ct[:2] = 0
    ct[H - 2:] = 0
    ct[H - 3, [W - 16, W - 15, W - 2, W - 1]] = 0
    return ct


def load_pinned():
    """Reference namespace whose rsgm._census_transform zeroes the unwritten census pixels (nothing else changes)."""
    ns = load()
    if not getattr(ns, "_pinned", False):
        orig = ns.rsgm._census_transform

        def _census_transform_pinned(left, right):
            ctl, ctr = orig(left, right)
            return zero_unwritten_census(ctl), zero_unwritten_census(ctr)

        ns.rsgm._census_transform = _census_transform_pinned
        ns._pinned = True
    return ns
