"""Fused TP (GEMM⊕reduce-scatter / all-gather⊕GEMM over NVLink) vs the NCCL baseline — needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [2, 4, 8])
def test_fused_tp_matches_nccl(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(29731 + n), os.path.join(root, "tests", "mp_tp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert "TP_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
