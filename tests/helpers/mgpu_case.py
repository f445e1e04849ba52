"""One small GRPO and one small PPO training step through the host API, on `world` ranks.
Used by tests/test_gpu_multirank.py: the single-process result is the reference, the torchrun
workers (one per GPU, NCCL) must reproduce it -- sharding whole groups per GPU changes nothing but
the fp32 summation order of the gradient."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def run_case(rank: int, world: int, groups: int = 8) -> dict:
    import trajopt_grpo_b200 as tg
    out = {}
    # ---- GRPO, Pendulum, 64x64 policy (tensor-core kernels)
    torch.manual_seed(0)
    pol = tg.GaussianActor_NeuralNetwork(3, 1, [64, 64], "ReLU", 0.5)
    opt = torch.optim.Adam(pol.parameters(), lr=1e-3)
    algo = tg.GRPO(0.2, 0.01, 0.99, pol, opt, None, updates_per_iter=3)
    mgr = tg.RolloutManager(lambda: tg.Pendulum(max_steps=40), pol, restart=True, num_workers=groups,
                            num_episodes_per_worker=16, use_multiprocessing=False, seed=3, rank=rank, world_size=world)
    buf = tg.Rollout_Buffer(mgr)
    buf.device_rollout = mgr.rollout_device()
    algo.learn(buf)
    out["grpo"] = pol.flat_parameters().detach().cpu().numpy().copy()
    out["_peer"] = algo._flat_opt._comm is not None
    # ---- PPO (GAE), QuadPole2D, 32x32 actor + critic (FP32-pipe kernels), full batch
    torch.manual_seed(1)
    pol2 = tg.GaussianActorCritic_NeuralNetwork(10, 2, [32, 32], "ReLU", 0.05)
    opt2 = torch.optim.Adam(pol2.parameters(), lr=1e-3)
    ppo = tg.PPO(0.2, pol2, opt2, None, 2, c1=0.5, kl_coeff=0.5, gamma=0.99, lam=0.95, entropy=0.01, batch_size=None,
                 monte_carlo=False)
    mgr2 = tg.RolloutManager(lambda: tg.QuadPole2D(max_steps=30), pol2, restart=False, num_workers=groups if groups != 8 else 6,
                             num_episodes_per_worker=8, use_multiprocessing=False, seed=5, rank=rank, world_size=world)
    buf2 = tg.Rollout_Buffer(mgr2)
    buf2.device_rollout = mgr2.rollout_device()
    ppo.learn(buf2)
    out["ppo"] = pol2.flat_parameters().detach().cpu().numpy().copy()
    return out


def tg_failure():
    from trajopt_grpo_b200 import engine
    return engine.PeerComm.last_failure


def main():
    import torch.distributed as dist
    ref_path = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    got = run_case(rank, world)
    ref = np.load(ref_path)
    if rank == 0:
        print("peer-memory allreduce:", got.pop("_peer"), "(", tg_failure(), ")")
    got.pop("_peer", None)
    for k, v in got.items():
        # identical on every rank (same allreduced gradient, same Adam)
        t = torch.from_numpy(v).cuda()
        all_t = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(all_t, t)
        for o in all_t:
            assert torch.equal(o, all_t[0]), f"{k}: weights differ between ranks"
        err = np.abs(v - ref[k]).max()
        assert err <= 2e-5 * max(1.0, np.abs(ref[k]).max()), (k, err)
        if rank == 0:
            print(f"{k}: max |sharded - single| = {err:.3e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
