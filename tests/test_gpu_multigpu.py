"""GPU suite, part 6: multi-GPU parity (needs >= 2 visible GPUs, skipped otherwise). Launches tests/multigpu_check.py under
torch.distributed.run: full-model ELBO and all parameter gradients of a batch sharded over 2 ranks (SyncBatchNorm statistics + one
flat gradient all-reduce) against the oracle on the GLOBAL batch (reference semantics: train.py:283, :314)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_training_step_matches_global_batch_oracle():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.environ.get('PYTHONPATH', ''))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
           '--master-port', str(29600 + os.getpid() % 300), os.path.join(ROOT, 'tests', 'multigpu_check.py')]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    print(r.stdout[-3000:])
    assert r.returncode == 0 and 'MULTIGPU CHECK PASS' in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
