"""z-slab domain decomposition: decomposed forward run over 2 GPUs (NCCL halo exchange) against the
single-GPU engine.  Needs two GPUs on the box; skipped otherwise (the host-side partition logic is covered on CPU in
tests/test_host_and_abi.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_slab_forward_matches_single_gpu_bitwise():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29571",
           os.path.join(root, "tests", "slab_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "bitwise equal = True" in r.stdout
