"""-m gpu tests that need >= 2 GPUs: the data-parallel engine on real NCCL against the single-device CPU oracle at the
GLOBAL batch (replicated tables, and row-sharded tables).  Skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _run(case, mode, nproc=2):
    if not torch.cuda.is_available() or torch.cuda.device_count() < nproc:
        pytest.skip(f"needs {nproc} GPUs")
    port = 29600 + (os.getpid() % 1000)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_worker.py"), case, mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0 and "multi_gpu_worker OK" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])


@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_data_parallel_engine_equals_oracle_at_global_batch(mode):
    """kkbox shape (BatchNorm on): 2 NCCL ranks x 32 samples vs the CPU oracle at 64 samples, 3 training steps."""
    _run("dp", mode)


@pytest.mark.parametrize("mode", ["fp32", "fp16"])
def test_row_sharded_tables_equal_oracle_at_global_batch(mode):
    """tmall shape, embedding / LR tables row-sharded over 2 ranks (NVLink peer loads in the gather, row gradients sent to
    their owners), vs the CPU oracle at the global batch; the gathered tables are compared after 3 steps."""
    _run("sharded", mode)
