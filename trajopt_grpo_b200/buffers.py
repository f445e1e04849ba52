"""Host-side mirror of buffers/rollout_buffer.py (and the empty tokenized_buffer.py).

`Rollout_Buffer` holds one rollout as CUDA tensors.  `sample()` keeps the data in
the kernels' struct-of-arrays layout (`self.device_rollout`) and exposes the
reference's attributes (`group_observations`, ... , rollout_buffer.py:55-70) as
strided views; `group_masks` is materialised lazily because the update kernels
work from the int32 lengths.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib as L
from .rollout import DeviceRollout


def _global_mean(x: torch.Tensor) -> float:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        s = torch.stack([x.sum(dtype=torch.float64), torch.tensor(float(x.numel()), dtype=torch.float64, device=x.device)])
        dist.all_reduce(s)
        return float((s[0] / s[1]).item())
    return float(x.mean().item())


class Buffer:
    """buffers/buffer.py:3-8."""

    def __init__(self):
        pass

    def store(self):
        pass


class Rollout_Buffer(Buffer):
    def __init__(self, rollout_manager, rtg: bool = True):
        self.rollout_manager = rollout_manager
        self.env = rollout_manager.env_fn()
        self.rtg = rtg
        self.group_observations = None
        self.group_actions = None
        self.group_rewards = None
        self.group_lengths = None
        self._group_masks = None
        self.device_rollout: DeviceRollout | None = None
        self.avg_reward = []

    # the reference's store() keeps the mask tensor; ours derives it from the lengths on first use
    @property
    def group_masks(self):
        if self._group_masks is None and self.device_rollout is not None:
            self._group_masks = self.device_rollout.group_masks()
        return self._group_masks

    @group_masks.setter
    def group_masks(self, value):
        self._group_masks = value

    def load(self, path: str):
        """rollout_buffer.py:31-42."""
        # atleast_1d: a one-epoch file makes np.loadtxt return a scalar (the reference then fails on len())
        self.avg_reward = np.atleast_1d(np.loadtxt(os.path.join(path, "reward.csv"), delimiter=",")).tolist()
        return len(self.avg_reward)

    def sample(self, init_state=None, noise=None, sync_metrics: bool = True):
        """rollout_buffer.py:45-53 through the fused kernel.  `init_state` ([S, local envs] CUDA tensor) and
        `noise` ([T, A, local envs]) are optional injection hooks (tests, benchmarks with host-provided
        initial states); by default the manager draws the reset distribution and the Philox stream.
        sync_metrics=False keeps the epoch free of host synchronisation: the mean return is appended as a
        0-dim CUDA tensor and converted when `avg_reward` is read through metadata() / save()."""
        r = self.rollout_manager.rollout_device(init_state=init_state, noise=noise)
        self.device_rollout = r
        self.group_observations = r.group_observations()
        self.group_actions = r.group_actions()
        self.group_rewards = r.group_rewards()
        self.group_lengths = r.group_lengths()
        self._group_masks = None
        # rollout_buffer.py:70: rewards.sum(2).mean() == mean episodic return (one scalar D2H); over ALL ranks'
        # groups when the rollout is sharded, so every rank logs (and rank 0 saves) the same history
        self.avg_reward.append(np.float32(_global_mean(r.ret)) if sync_metrics else r.ret.mean())

    def store(self, group_observations, group_actions, group_rewards, group_lengths, group_masks):
        """rollout_buffer.py:55-70 for externally produced [G,E,T,.] tensors: they are
        re-laid out once into the kernels' struct-of-arrays form."""
        dev = torch.device("cuda", torch.cuda.current_device())
        obs = torch.as_tensor(group_observations, dtype=torch.float32).to(dev)
        act = torch.as_tensor(group_actions, dtype=torch.float32).to(dev)
        rew = torch.as_tensor(group_rewards, dtype=torch.float32).to(dev)
        G, E, T, _ = obs.shape
        N = G * E
        ln = torch.as_tensor(group_lengths).to(dev).reshape(N).to(torch.int32)
        soa = DeviceRollout(
            obs=obs.reshape(N, T, -1).permute(1, 2, 0).contiguous(),
            act=act.reshape(N, T, -1).permute(1, 2, 0).contiguous(),
            rew=rew.reshape(N, T).t().contiguous(),
            logp=None, len=ln.contiguous(), ret=rew.reshape(N, T).sum(1), G=G, E=E, T=T)
        self.device_rollout = soa
        self.group_observations, self.group_actions = soa.group_observations(), soa.group_actions()
        self.group_rewards, self.group_lengths = soa.group_rewards(), soa.group_lengths()
        self._group_masks = torch.as_tensor(group_masks, dtype=torch.float32).to(dev)
        self.avg_reward.append(rew.sum(2).mean().detach().cpu().numpy())

    def save_trajectory(self, path: str):
        """rollout_buffer.py:72-102: CSV of (episode_id, observation_i..., action_i...) for valid steps.
        The valid rows are compacted on the device (tg_export_trajectory) and copied to the host
        once through pinned memory; the reference's float64 table is reproduced for the CSV."""
        import pandas as pd
        from . import engine
        r = self.device_rollout
        if r is None:
            raise L.EngineError("buffer holds no rollout: call sample() or store(...) first")
        ids_d, rows_d = engine.export_trajectory(r.obs, r.act, r.len)
        ids = torch.empty(ids_d.shape, dtype=ids_d.dtype).pin_memory()
        rows = torch.empty(rows_d.shape, dtype=rows_d.dtype).pin_memory()
        ids.copy_(ids_d, non_blocking=True)
        rows.copy_(rows_d, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        O, A = r.obs.shape[1], r.act.shape[1]
        header = ["episode_id"] + [f"observation_{i}" for i in range(O)] + [f"action_{i}" for i in range(A)]
        data = np.hstack([ids.numpy().reshape(-1, 1).astype(np.float64), rows.numpy().astype(np.float64)])
        df = pd.DataFrame(data, columns=header)
        df["episode_id"] = df["episode_id"].astype(int)
        df.to_csv(os.path.join(path, "trajectory.csv"), index=False)

    def host_slice(self, num_groups: int = 4, num_episodes: int = 5):
        """The Dashboard / Publisher feed (visualize/visualizer.py:105-142, publish/publisher.py:38-64 draw
        the first `max_episodes_per_render` episodes of the first 4 groups): copies ONLY that slice to the
        host -- [g, e, T, .] numpy arrays plus integer lengths -- so rendering never moves the full trajectory."""
        r = self.device_rollout
        if r is None:
            raise L.EngineError("buffer holds no rollout: call sample() or store(...) first")
        g, e = min(num_groups, r.G), min(num_episodes, r.E)
        idx = (torch.arange(g, device=r.obs.device)[:, None] * r.E + torch.arange(e, device=r.obs.device)[None, :]).reshape(-1)
        obs = r.obs.index_select(2, idx).permute(2, 0, 1).reshape(g, e, r.T, -1)
        act = r.act.index_select(2, idx).permute(2, 0, 1).reshape(g, e, r.T, -1)
        rew = r.rew.index_select(1, idx).t().reshape(g, e, r.T)
        ln = r.len.index_select(0, idx).reshape(g, e)
        return {"observations": obs.cpu().numpy(), "actions": act.cpu().numpy(), "rewards": rew.cpu().numpy(),
                "lengths": ln.cpu().numpy().astype(int)}

    def _resolve_rewards(self):
        """deferred (sync_metrics=False) entries -> floats."""
        for i, v in enumerate(self.avg_reward):
            if isinstance(v, torch.Tensor):
                self.avg_reward[i] = np.float32(v.item())

    def metadata(self):
        self._resolve_rewards()
        return {"avg_reward": float(self.avg_reward[-1]) if len(self.avg_reward) > 0 else None}

    def save(self, path: str):
        """rollout_buffer.py:115-126."""
        self._resolve_rewards()
        if not L.is_checkpoint_writer():
            return
        with open(os.path.join(path, "reward.csv"), "w") as f:
            for reward in self.avg_reward:
                f.write(f"{reward}\n")


class TokenizedBuffer(Rollout_Buffer):
    """buffers/tokenized_buffer.py is an EMPTY file in the reference (the name only
    appears in README.md:53); kept as an alias so imports resolve."""
