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

from .rollout import DeviceRollout


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

    def sample(self):
        """rollout_buffer.py:45-53 through the fused kernel."""
        r = self.rollout_manager.rollout_device()
        self.device_rollout = r
        self.group_observations = r.group_observations()
        self.group_actions = r.group_actions()
        self.group_rewards = r.group_rewards()
        self.group_lengths = r.group_lengths()
        self._group_masks = None
        # rollout_buffer.py:70: rewards.sum(2).mean() == mean episodic return (one scalar D2H)
        self.avg_reward.append(np.float32(r.ret.mean().item()))

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
        """rollout_buffer.py:72-102: CSV of (episode_id, observation_i..., action_i...) for valid steps."""
        import pandas as pd
        obs = self.group_observations.detach().cpu().numpy()
        act = self.group_actions.detach().cpu().numpy()
        ln = self.group_lengths.detach().cpu().numpy().astype(int)
        rows_o, rows_a, ids = [], [], []
        for i in range(ln.shape[0]):
            for j in range(ln.shape[1]):
                rows_o.append(obs[i, j, :ln[i, j]])
                rows_a.append(act[i, j, :ln[i, j]])
                ids.extend([j + i * ln.shape[1]] * ln[i, j])
        header = ["episode_id"] + [f"observation_{i}" for i in range(obs.shape[3])] + \
                 [f"action_{i}" for i in range(act.shape[3])]
        data = np.hstack([np.array(ids).reshape(-1, 1), np.vstack(rows_o), np.vstack(rows_a)])
        df = pd.DataFrame(data, columns=header)
        df["episode_id"] = df["episode_id"].astype(int)
        df.to_csv(os.path.join(path, "trajectory.csv"), index=False)

    def metadata(self):
        return {"avg_reward": float(self.avg_reward[-1]) if len(self.avg_reward) > 0 else None}

    def save(self, path: str):
        """rollout_buffer.py:115-126."""
        with open(os.path.join(path, "reward.csv"), "w") as f:
            for reward in self.avg_reward:
                f.write(f"{reward}\n")


class TokenizedBuffer(Rollout_Buffer):
    """buffers/tokenized_buffer.py is an EMPTY file in the reference (the name only
    appears in README.md:53); kept as an alias so imports resolve."""
