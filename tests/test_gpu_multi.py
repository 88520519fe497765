"""GPU, 2+ devices: N slabs == 1 GPU for the PRODUCT (vsb_step + peer-memory halo kernels + NCCL fallback), as a test.

Launches tests/multi_gpu_worker.py under torch.distributed.run on min(device_count, 4) GPUs: every rank advances its
slab, rank 0 also advances the whole domain on one GPU; periodic KBC, the C2 recipe with walls and an immersed
cylinder, D3Q19 BGK and D3Q19 MRT with a densely meshed body, each through the NCCL exchange and through the
peer-memory exchange inside a CUDA graph.  Skipped on a single-GPU box (the CPU suite covers the exchange indices
with gloo; bench.py reports `parity_vs_1gpu` on every multi-GPU line)."""

import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_n_slabs_equal_one_gpu():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
