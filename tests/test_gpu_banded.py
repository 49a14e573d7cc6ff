"""A frame split into row bands: the aggregation of each band continues from the row state of its
neighbour, so the banded compute_rsgm equals the unsplit one BIT FOR BIT (unlike the reference's approximate StripedStereoSGM,
RSGM/StereoSGM.h:116-133).  One GPU: bands one after the other; two GPUs: one band per rank, state and disparities over NVLink."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, assert_same

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,D,nb", [((40, 120), 64, 2), ((61, 200), 192, 3), ((90, 330), 48, 5), ((33, 70), 32, 1), ((128, 700), 96, 4),
                                        ((50, 100), 8, 7)])
def test_bands_on_one_gpu_equal_unsplit(orc, shape, D, nb):
    from vppstereo_b200 import rsgm, synth
    from vppstereo_b200.banded import BandedRsgm
    p = synth.make_pair(700 + D + nb, shape=shape, hints="random")
    want = rsgm.compute_rsgm(p["left"], p["left"], p["right"], dmax=D)
    ref = orc.compute_rsgm(p["left"], p["left"], p["right"], dmax=D)
    assert_same(want, ref, "unsplit vs oracle")
    b = BandedRsgm(shape[0], shape[1], 3, dmax=D, n_bands=nb)
    assert len(b.bands) == min(nb, b.Hp // 3) and sum(r for _, r in b.bands) == b.Hp
    got = b.compute(p["left"], p["left"], p["right"])
    assert_same(got, want, f"{nb} bands {shape} D={D}")
    got2 = b.compute(p["left"], p["left"], p["right"])                 # buffers reused
    assert_same(got2, want, "second call")


def test_bands_wide_strips_and_gray(orc, monkeypatch):
    """teams of several CTAs inside a band (forced narrow strips), gray input, sub-pixel off"""
    from vppstereo_b200 import _lib, rsgm, synth
    from vppstereo_b200.banded import BandedRsgm
    p = synth.make_pair(31, shape=(70, 400), hints="random", channels=1)
    for strip in (0, 64):
        _lib.set_tuning(_lib.TUNE_SGM_MAX_STRIP, strip)
        try:
            want = rsgm.compute_rsgm(p["left"], p["left"], p["right"], dmax=64, subpixel=False)
            got = BandedRsgm(70, 400, 1, dmax=64, n_bands=3, subpixel=False).compute(p["left"], p["left"], p["right"])
            assert_same(got, want, f"strip {strip}")
        finally:
            _lib.set_tuning(_lib.TUNE_SGM_MAX_STRIP, 0)


def test_one_band_per_rank_two_gpus(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29900 + os.getpid() % 90
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_banded_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for rank in range(2):
        text = open(tmp_path / f"rank{rank}.txt").read()
        assert text.startswith("OK"), text
