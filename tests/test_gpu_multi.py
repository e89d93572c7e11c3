"""Multi-rank parity.  Row-sharded A (fused peer-memory exchange and NCCL all-reduce) reproduces the single-GPU
iterates; a batch split across ranks reproduces the single batch bit for bit.  Launches scripts/multi_gpu_check.py
and scripts/multi_gpu_batch_check.py under torchrun with two ranks: one rank per GPU when the box has two GPUs,
otherwise BOTH RANKS ON THE ONE GPU (two processes, CUDA-IPC peer exchange between them, kernels time-sliced), so
that the multi-rank code paths run -- not skip -- on a single-GPU test box."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _torchrun(script, port, extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), str(ROOT / "scripts" / script)] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0


def _extra():
    import torch
    return [] if torch.cuda.device_count() >= 2 else ["--same-device"]


def test_row_sharded_two_ranks():
    _torchrun("multi_gpu_check.py", 29517, _extra())


def test_batch_split_two_ranks():
    _torchrun("multi_gpu_batch_check.py", 29518, _extra())
