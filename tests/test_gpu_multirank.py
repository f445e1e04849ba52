"""Two ranks over NCCL (one process per GPU, whole groups per GPU) against the single-GPU run of the
same GRPO and PPO steps.  Needs two visible GPUs; skipped otherwise (the world-size-2 host logic is
covered on CPU by tests/test_multirank_gloo.py, the sharded PPO arithmetic on one GPU by
tests/test_gpu_kernels.py::test_ppo_sharded_equals_single_rollout)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HELPER = os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers", "mgpu_case.py")


@pytest.mark.parametrize("peer", ["1", "0"])
def test_two_ranks_reproduce_single_gpu_training_step(tmp_path, peer):
    """peer = "1": gradient allreduce + Adam fused over NVLink peer memory (tg_allreduce_adam_step, the default);
    peer = "0": the NCCL allreduce + tg_adam_step fallback.  Both must reproduce the single-GPU weights and leave
    bit-identical weights on the two ranks."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.dirname(HELPER))
    import mgpu_case
    ref = mgpu_case.run_case(0, 1)
    ref_path = str(tmp_path / "ref.npz")
    np.savez(ref_path, **ref)
    env = dict(os.environ, TG_PEER_ALLREDUCE=peer)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533" if peer == "1" else "29534", HELPER, ref_path],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert ("peer-memory allreduce: True" in r.stdout) == (peer == "1"), r.stdout[-2000:]
