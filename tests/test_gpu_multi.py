"""Multi-GPU parity (needs >= 2 GPUs on the box): row-sharded A with the NCCL all-reduce inside the
library reproduces the single-GPU iterates.  Launches scripts/multi_gpu_check.py under torchrun."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def test_row_sharded_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", str(ROOT / "scripts" / "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0
