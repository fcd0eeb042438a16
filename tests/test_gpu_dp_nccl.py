"""Data-parallel equivalence ON HARDWARE (SURVEY §4 "Distributed" row; train_ddp.py:79): needs >= 2 GPUs on the box, one
process per GPU over NCCL (skipped on single-GPU boxes; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp_nccl.py -m gpu`).
The CPU suite covers the same host logic over gloo (tests/test_dp_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_allreduced_gradients_equal_single_process_union_and_replicas_stay_identical():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (one NCCL rank per GPU)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dp_nccl_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0 and "DP_NCCL_OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
