"""Multi-GPU tests (skipped on a single-GPU box): the data-parallel training path over NCCL and over NVLink peer memory."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize('shape', ['mma', 'tc'])
def test_dp_p2p_fused_allreduce_matches_nccl(shape):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', '29533', os.path.join(ROOT, 'tests', 'dp_p2p_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, DP_CHECK_SHAPE=shape))
    assert r.returncode == 0 and 'DP_P2P_CHECK_OK' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
