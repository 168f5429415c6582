"""N > 1 on real GPUs: torchrun launches tools/multigpu_check.py (NCCL) when the box has at least two GPUs.
The host-side logic of the same module is covered on CPU by tests/test_distributed_cpu.py (gloo)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_member_and_sequence_sharding_on_two_gpus(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run under gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29511", os.path.join(ROOT, "tools", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert '"ok": true' in r.stdout


def test_sequence_sharding_is_independent_of_the_split_single_gpu(cuda):
    """seq_base keys the Philox stream: sequences 2..3 decoded alone equal rows 2..3 of the batch of 4."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "multigpu_check.py")], capture_output=True,
                       text=True, timeout=900, env={**os.environ, "WORLD_SIZE": "1", "RANK": "0", "LOCAL_RANK": "0"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert '"sequence_sharding_tokens_equal": true' in r.stdout
