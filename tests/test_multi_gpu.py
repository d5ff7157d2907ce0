"""Multi-GPU parity on hardware (needs >= 2 GPUs in the box, skipped otherwise): tools/dist_check.py under torchrun --
distributed vmult / CG / 3-component CG with hanging nodes against the same problem on one GPU, once per transport
(P2P peer stores over CUDA-IPC windows, NCCL send/recv), component-batched vs per-component exchange."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dist_check_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "tools", "dist_check.py")], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "FAIL" not in r.stdout and "DIST_CHECK_HANGING [nccl] PASS" in r.stdout
