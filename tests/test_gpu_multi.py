"""The data-parallel path on real GPUs: runs scripts/multi_gpu_check.py under torchrun on every visible
GPU — peer all-reduce vs NCCL, N-rank training step / global-batch memory / sharded scoring vs the single-process
result.  Skipped on a one-GPU box (the driver's GPU test tier); run with ``gpurun --gpus 2``."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_data_parallel_path_on_real_gpus():
    n = min(torch.cuda.device_count(), 8)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "scripts", "multi_gpu_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI-GPU CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
