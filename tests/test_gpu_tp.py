"""Tensor-parallel linear on two real GPUs (row (e) of SURVEY §8): NCCL all-gather and the
gather fused into the GEMM epilogue both equal the single-device layer bit for bit.
Skipped on a one-GPU box; the host-side logic is covered on CPU by test_sharding_cpu.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_tensor_parallel_linear_two_gpus():
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_tp_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", worker]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0 and "TP_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
