"""Tile sharding over >= 2 GPUs (NCCL): skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tile_sharded_forward_matches_single_gpu(cuda):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs')
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr',
           '127.0.0.1', '--master-port', '29641', os.path.join(ROOT, 'tools', 'check_shard.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=800)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0
