"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): runs tools/mgpu_check.py under
torchrun on all visible GPUs (capped at 8)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_union_over_ranks_equals_oracle():
    import torch
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(ROOT, "tools", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=900)
    assert "MGPU_CHECK_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
