"""GPU test of the fold-sharded multi-GPU path over NCCL (needs >= 2 GPUs; the single-GPU driver run skips it).
scripts/dist_check.py shards the folds of one utterance over the ranks and compares with the single-GPU waveform
bit for bit (the single-GPU waveform itself is checked against the oracle in test_gpu_parity.py)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two CUDA devices")
def test_sharded_generation_matches_single_gpu_over_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "dist_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("bit-identical to single GPU: True") == 4, res.stdout[-2000:]
