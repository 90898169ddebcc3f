"""GPU, >= 2 devices: the real NCCL exchange path (one process per GPU) against the unsharded oracle."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_real_ranks(p2p):
    """p2p = 1: the exchanges run inside the kernels over the peer-memory mailboxes (ParticleSwarm step, every stage of the
    CrossEntropy pass); p2p = 0 (EQS_P2P=0): the same passes phase by phase with ncclAllGather between the kernels."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (virtual-rank tests cover the protocol on one GPU)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517" if p2p == "1" else "29518",
           os.path.join(ROOT, "tests", "multigpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=dict(os.environ, EQS_P2P=p2p))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "multigpu_check OK" in r.stdout
