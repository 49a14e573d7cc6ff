"""Two GPUs of one box, one process per GPU over NCCL (torch.distributed.run, rendezvous on 127.0.0.1): the frame-sharded path on
real kernels.  Gathered disparities (copy-engine PeerGather and the NCCL all_gather fallback) == what a single GPU computes
for the same frames.  Skipped on a single-GPU box; tests/test_dist_gloo.py covers the host logic on CPU."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_gathered_equals_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29700 + os.getpid() % 200
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for rank in range(2):
        text = open(tmp_path / f"rank{rank}.txt").read()
        assert text.startswith("OK"), text
        assert "p2p: available=True" in text, text          # the copy-engine path really ran
