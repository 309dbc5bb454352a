"""Marker-sharded multi-GPU parity (NCCL allreduce of X.v, fused scalar allreduces); needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_parity():
    from gvamp_b200 import capi
    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port",
           "29511", os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "mgpu_check ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
