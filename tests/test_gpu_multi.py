"""Sharded solve on several GPUs of one node (needs >= 2 visible GPUs: skipped on a single-GPU box).  Spawns
tools/multigpu_check.py under torchrun: landmarks dealt over the ranks, the reduced system exchanged per GN iteration
(ncclAllReduce, then the same handle with the NVLink mailbox exchange), sharded == unsharded for the linearised system,
the LM solve, and updatePointStatuses (cross-rank radix select); replicas bitwise identical."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("peer", [0, 1, 2])
def test_sharded_solve_equals_unsharded_on_all_visible_gpus(peer):
    n = min(_gpus(), 8)
    if n < 2:
        pytest.skip("one GPU visible: the multi-GPU path is covered by tests/test_sharding_gloo.py (CPU, world_size 2) "
                    "and by tools/gpu_multi.sh on multi-GPU boxes")
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    env = dict(os.environ, DPBA_SPEC_MULTI="1")
    if peer:
        env["DPBA_PEER_EXCHANGE"] = "1"  # 1: stand-alone mailbox kernel, scalars and system in two concurrent exchanges (the default)
        env["DPBA_PEER_FUSED"] = "1" if peer == 2 else "0"  # 2: exchange fused into the producers / consumers
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + peer), os.path.join(ROOT, "tools", "multigpu_check.py")]
    run = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=400)
    assert run.returncode == 0 and f"MULTIGPU_CHECK PASS world={n}" in run.stdout, run.stdout[-3000:] + run.stderr[-3000:]
