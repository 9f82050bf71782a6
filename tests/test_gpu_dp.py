"""Data parallelism on real GPUs (needs >= 2): tests/dp_worker.py under torchrun, with the one-shot peer-memory exchanges
and with NCCL only. Skipped on a single-GPU box (run with `gpurun --gpus 2`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("peer", ["1", "0"])
def test_two_rank_data_parallel_matches_global_batch_oracle(peer):
    n = _ngpus()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    env = dict(os.environ, B200NN_PEER=peer)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517" if peer == "1" else "29518", os.path.join(ROOT, "tests", "dp_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "DP_OK" in r.stdout, r.stdout[-3000:] + "\n" + r.stderr[-6000:]
