"""GPU, >= 2 devices: domain-decomposed run (one process per GPU, NCCL halo exchange of state and gradients
each RK stage) against the CPU oracle.  Skipped on a single-GPU box; the CPU suite covers the partition and
halo plan with gloo (tests/test_gloo_halo.py)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_two_rank_parity():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "scripts", "mgpu_parity.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert "PARITY_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
