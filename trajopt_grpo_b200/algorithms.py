"""Host-side mirror of algorithms/algorithm.py, grpo.py and ppo.py.

Same constructors, `learn(buffer)`, `save/load/metadata` and `old_policy`
attribute as the reference; `learn` is a short sequence of kernel launches:

  GRPO.learn (grpo.py:50-148):
      tg_advantage(GRPO)                         RTG + per-group z-score
      [tg_policy_forward_traj]                   frozen old-policy log-prob (skipped when
                                                 the rollout's own log-prob is that value)
      updates_per_iter x { tg_policy_grad  ->  [NCCL allreduce]  ->  tg_adam_step }

The objective keeps the reference's sign: `J.backward(); optimizer.step()` with a
minimising optimizer descends on J -- pass `maximize=True` to ascend.
"""
from __future__ import annotations

import copy
import os
from abc import ABC, abstractmethod

import torch

from . import _lib as L
from . import engine


class Algorithm(ABC):
    """algorithms/algorithm.py:3-36."""

    def __init__(self):
        pass

    @abstractmethod
    def learn(self):
        pass

    @abstractmethod
    def metadata(self):
        return {}

    @abstractmethod
    def save(self, path: str):
        pass

    @abstractmethod
    def load(self, path: str):
        pass


class _FlatOptimizer:
    """Applies a flat gradient to the policy's flat parameter vector.

    torch.optim.Adam with default flags (what every reference pipeline builds,
    pipelines/cartpole_pipeline_grpo.py:65) runs as ONE tg_adam_step launch whose
    moment buffers are aliased into `optimizer.state`, so `optimizer.state_dict()`
    checkpoints (grpo.py:150-160, ppo.py:207-225) keep working and
    `optimizer.load_state_dict()` (a resumed run) is picked up: state tensors that are
    not our aliases were put there by torch, so their moments are copied in and their
    step count is taken verbatim.  Any other optimizer gets `p.grad` views of the flat
    gradient and its own `step()`.
    """

    def __init__(self, optimizer, policy):
        self.opt, self.policy = optimizer, policy
        self.m = self.v = None
        self.step_count = 0
        self._steps = {}          # id(param) -> its own CPU step tensor (torch's Adam keeps one per parameter)
        self._ok_key = None
        self._ok = False
        self._comm = None         # engine.PeerComm: gradient allreduce + Adam over NVLink peer memory (world > 1)
        self._comm_tried = False

    def _fused_adam_ok(self, flat):
        groups = self.opt.param_groups
        g = groups[0] if groups else {}
        key = (type(self.opt), len(groups), id(g.get("params")), len(g.get("params", ())), flat.data_ptr(),
               flat.numel(), g.get("amsgrad"), g.get("maximize"), g.get("weight_decay"), g.get("capturable"),
               g.get("differentiable"))
        if key == self._ok_key:
            return self._ok
        ok = type(self.opt) is torch.optim.Adam and len(groups) == 1
        if ok:
            off, base = 0, flat.data_ptr()
            for p in g["params"]:                                     # parameters alias the flat vector, in order
                if p.data_ptr() != base + 4 * off:
                    ok = False
                    break
                off += p.numel()
            ok = ok and off == flat.numel()
        ok = ok and not g.get("amsgrad", False) and not g.get("maximize", False) and g.get("weight_decay", 0) == 0 \
            and not g.get("capturable", False) and not g.get("differentiable", False)
        self._ok_key, self._ok = key, ok
        return ok

    def _sync_adam_state(self, flat):
        if self.m is None or self.m.numel() != flat.numel() or self.m.device != flat.device:
            self.m, self.v = torch.zeros_like(flat), torch.zeros_like(flat)
            self._steps = {}
        off = 0
        loaded_step = None
        for p in self.opt.param_groups[0]["params"]:
            n = p.numel()
            st = self.opt.state[p]
            for key, buf in (("exp_avg", self.m), ("exp_avg_sq", self.v)):
                view = buf[off:off + n].view(p.shape)
                cur = st.get(key)
                if cur is not None and cur.data_ptr() != view.data_ptr():
                    view.copy_(cur.to(view.device, torch.float32))      # state came from load_state_dict()
                st[key] = view
            mine = self._steps.get(id(p))
            cur = st.get("step")
            if cur is not None and cur is not mine:                     # replaced by load_state_dict(): verbatim
                loaded_step = int(float(cur))
            if mine is None:
                mine = self._steps[id(p)] = torch.tensor(float(self.step_count))
            st["step"] = mine
            off += n
        if loaded_step is not None:
            self.step_count = loaded_step
            for t in self._steps.values():
                t.fill_(float(loaded_step))

    def peer_comm(self, flat):
        """The peer-memory gradient window of a sharded run (built on first use, collectively), or None: world
        size 1, an optimizer the fused Adam kernel does not cover, TG_PEER_ALLREDUCE=0, or the windows could
        not be mapped on every rank (the NCCL allreduce + tg_adam_step path is then used)."""
        if self._comm_tried:
            return self._comm if (self._comm is None or self._comm.n >= flat.numel()) else None
        self._comm_tried = True
        if _dist_world() > 1 and self._fused_adam_ok(flat) and os.environ.get("TG_PEER_ALLREDUCE", "1") != "0":
            self._comm = engine.PeerComm.try_create(flat.numel(), flat.device)
        return self._comm

    def step_allreduced(self, flat, comm):
        """Adam on the sum over ranks of the gradients the ranks wrote into their comm slots (one fused launch)."""
        self._sync_adam_state(flat)
        g = self.opt.param_groups[0]
        self.step_count += 1
        comm.allreduce_adam_step(flat, self.m, self.v, self.step_count, g["lr"], g["betas"][0], g["betas"][1], g["eps"])
        for t in self._steps.values():
            t.fill_(float(self.step_count))

    def step(self, flat, grad):
        if self._fused_adam_ok(flat):
            self._sync_adam_state(flat)
            g = self.opt.param_groups[0]
            self.step_count += 1
            engine.adam_step(flat, grad, self.m, self.v, self.step_count, g["lr"], g["betas"][0], g["betas"][1],
                             g["eps"])
            for t in self._steps.values():
                t.fill_(float(self.step_count))
            return
        off = 0
        params = [p for grp in self.opt.param_groups for p in grp["params"]]
        if sum(p.numel() for p in params) != grad.numel():
            raise L.EngineError("optimizer parameters do not match the policy's flat parameter vector")
        for p in params:
            p.grad = grad[off:off + p.numel()].view(p.shape)
            off += p.numel()
        self.opt.step()


_WORLD_OVERRIDE = None


class single_process:
    """with algorithms.single_process(): learn() treats the rollout as unsharded even though a process group is
    initialised (used to compute the single-GPU reference of a sharded run inside the same process)."""

    def __enter__(self):
        global _WORLD_OVERRIDE
        self._prev, _WORLD_OVERRIDE = _WORLD_OVERRIDE, 1
        return self

    def __exit__(self, *exc):
        global _WORLD_OVERRIDE
        _WORLD_OVERRIDE = self._prev
        return False


def _dist_world():
    if _WORLD_OVERRIDE is not None:
        return _WORLD_OVERRIDE
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return 1


def _rollout_of(buffer):
    r = getattr(buffer, "device_rollout", None)
    if r is None:
        if getattr(buffer, "group_observations", None) is None:
            raise L.EngineError("buffer holds no rollout: call buffer.sample() or buffer.store(...) first")
        from .buffers import Rollout_Buffer
        tmp = Rollout_Buffer.__new__(Rollout_Buffer)
        tmp.avg_reward = []
        Rollout_Buffer.store(tmp, buffer.group_observations, buffer.group_actions, buffer.group_rewards,
                             buffer.group_masks.sum(-1) if getattr(buffer, "group_lengths", None) is None
                             else buffer.group_lengths, buffer.group_masks)
        r = tmp.device_rollout
    return r


class GRPO(Algorithm):
    """algorithms/grpo.py:12-168."""

    def __init__(self, epsilon: float, beta: float, gamma: float, policy, optimizer, ref_model=None,
                 updates_per_iter: int = 10, *, maximize: bool = False):
        self.epsilon, self.policy, self.ref_model, self.beta = epsilon, policy, ref_model, beta
        self.gamma, self.updates_per_iter, self.optimizer = gamma, updates_per_iter, optimizer
        self.maximize = maximize
        if ref_model is not None:
            # grpo.py:129-132 cannot run in the reference either; no oracle for a KL term
            raise L.EngineError("ref_model is not supported: the reference's KL branch is broken (every config passes None)")
        self.old_policy = copy.deepcopy(self.policy)                 # grpo.py:48
        self._synced_tag = policy.param_tag()                        # old_policy == policy at this tag ...
        self._old_tag = self.old_policy.param_tag()                  # ... while old_policy itself is untouched
        self._seen_loads = getattr(policy, "checkpoint_loads", 0)
        self._flat_opt = _FlatOptimizer(optimizer, policy)
        self.last_stats = None

    def learn(self, buffer) -> None:
        r = _rollout_of(buffer)
        pol = self.policy
        self._resync_after_checkpoint_load()
        flat = pol.flat_parameters()
        old_flat = self.old_policy.flat_parameters()
        dims, act_name, cov = pol.actor.dims, pol.actor.activation_name, pol.cov_diag
        world = _dist_world()
        G_global = r.G * world
        # grpo.py:66-74, 108-115
        adv, _ = engine.advantage(L.ADV_GRPO, r.G, r.E, r.T, self.gamma, 0.0, r.rew, r.len)
        # grpo.py:118-119: log-prob under the frozen old policy.  The rollout kernel already
        # evaluated it when the rollout was produced by weights identical to old_policy.
        cur_tag = pol.param_tag()
        if r.logp is not None and r.policy_tag == cur_tag and self._synced_tag == cur_tag \
                and self.old_policy.param_tag() == self._old_tag:
            old_logp = r.logp
        else:
            _, old_logp = engine.policy_forward_traj(dims, act_name, old_flat[:flat.numel()].contiguous(), r.obs, cov,
                                                     r.act, r.len)
        scale = (-1.0 if self.maximize else 1.0) / G_global         # grpo.py:140 (J /= group_size)
        comm = self._flat_opt.peer_comm(flat) if world > 1 else None
        with engine.length_order(r.len, r.T):                       # lengths are fixed for all updates of this learn()
            for _ in range(self.updates_per_iter):                  # grpo.py:106
                # the path's only collective: the gradient summed over ranks.  With the peer-memory window the kernel
                # writes its gradient into this rank's slot and ONE launch does allreduce + Adam (tg_comm.cu);
                # otherwise NCCL allreduce + tg_adam_step
                slot = comm.grad_slot()[:flat.numel()] if comm is not None else None
                grad, stats = engine.policy_grad(dims, act_name, flat, cov, r.obs, r.act, adv, old_logp, r.len,
                                                 self.epsilon, scale, out_grad=slot)
                if comm is not None:
                    self._flat_opt.step_allreduced(flat, comm)      # grpo.py:143-145 on the summed gradient
                else:
                    if world > 1:
                        import torch.distributed as dist
                        dist.all_reduce(grad)
                    self._flat_opt.step(flat, grad)                 # grpo.py:143-145
                pol.bump_param_epoch()
                self.last_stats = stats
        self.old_policy.load_state_dict(self.policy.state_dict())   # grpo.py:148
        self._synced_tag = pol.param_tag()
        self._old_tag = self.old_policy.param_tag()

    def learn_streamed(self, manager, init_state=None, chunk_groups: int = 8192):
        """One whole epoch -- rollout AND learn (pipelines/pipeline.py:163-164) -- for more envs than trajectories fit
        in HBM (BASELINE configs[4]: up to 4M envs per GPU x 1000 steps x 104 B = 416 GB).

        The reference stores every trajectory (rollout_worker.py:64-68) because its GRPO.learn revisits them in
        every update (grpo.py:106-145).  Here only the initial states and the Philox seed are kept; the rollout is
        REMATERIALISED chunk by chunk (whole GRPO groups, so the group statistics of grpo.py:108-115 stay exact):

            for each update:  for each chunk of `chunk_groups` groups:
                K1 rollout of the chunk under the FROZEN old policy (deterministic: same seed, same global env
                   index -> the identical trajectories, log-probs included, in every update)
                K2 advantages of the chunk         K3 gradient of the chunk under the current weights, accumulated
            [gradient allreduce]  Adam

        One trajectory buffer of chunk size is reused.  With updates_per_iter == 1 no work is repeated.
        Equal to sample() + learn() on the same seed up to the fp32 summation order of the per-chunk gradients.
        Returns the number of valid env-steps of this rank's rollout as a 0-dim CUDA tensor (no host sync)."""
        from .rollout import shard_initial_states
        import numpy as np
        pol = self.policy
        self._resync_after_checkpoint_load()
        flat = pol.flat_parameters()
        if self.old_policy.param_tag() != self._old_tag or self._synced_tag != pol.param_tag():
            self.old_policy.load_state_dict(pol.state_dict())      # start of an epoch: old == current (grpo.py:148)
        old_flat = self.old_policy.actor.flat_params()
        dims, act_name, cov = pol.actor.dims, pol.actor.activation_name, pol.cov_diag
        env, E = manager.env, manager.num_episodes_per_worker
        G_loc = manager.local_workers
        world = _dist_world()
        dev = flat.device
        if init_state is None:
            blk = shard_initial_states(env, manager.num_workers, E, manager.restart, manager._rng, manager.rank,
                                       manager.world_size)
            init_state = torch.from_numpy(np.ascontiguousarray(blk.T)).to(torch.float32).pin_memory().to(dev, non_blocking=True)
        seed = (manager._seed + 0x9E3779B97F4A7C15 * (manager._epoch + 1)) & (2 ** 64 - 1)
        manager._epoch += 1
        scale = (-1.0 if self.maximize else 1.0) / (G_loc * world)
        n_valid = torch.zeros((), dtype=torch.int64, device=dev)
        ret_sum = torch.zeros((), dtype=torch.float64, device=dev)
        out = getattr(self, "_stream_out", None)
        grad_own, grad_part = torch.empty_like(flat), torch.empty_like(flat)
        comm = self._flat_opt.peer_comm(flat) if world > 1 else None
        for u in range(self.updates_per_iter):
            grad_total = comm.grad_slot()[:flat.numel()] if comm is not None else grad_own
            grad_total.zero_()
            for g0 in range(0, G_loc, chunk_groups):
                g1 = min(G_loc, g0 + chunk_groups)
                n0, n1 = g0 * E, g1 * E
                s0 = init_state[:, n0:n1].contiguous()
                out = engine.rollout(env._tg_kind, env.max_steps, env.timestep, dims, act_name, old_flat, cov, s0,
                                     seed=seed, env_offset=manager.env_offset + n0, phys=getattr(env, "_tg_phys", None),
                                     out=out)
                adv, _ = engine.advantage(L.ADV_GRPO, g1 - g0, E, env.max_steps, self.gamma, 0.0, out["rew"], out["len"])
                _, stats = engine.policy_grad(dims, act_name, flat, cov, out["obs"], out["act"], adv, out["logp"],
                                              out["len"], self.epsilon, scale, out_grad=grad_part)
                grad_total.add_(grad_part)
                if u == 0:
                    n_valid.add_(out["len"].sum())
                    ret_sum.add_(out["ret"].sum(dtype=torch.float64))
                self.last_stats = stats
            if comm is not None:
                self._flat_opt.step_allreduced(flat, comm)
            else:
                if world > 1:
                    import torch.distributed as dist
                    dist.all_reduce(grad_total)
                self._flat_opt.step(flat, grad_total)
            pol.bump_param_epoch()
        self._stream_out = out
        self.old_policy.load_state_dict(self.policy.state_dict())   # grpo.py:148
        self._synced_tag = pol.param_tag()
        self._old_tag = self.old_policy.param_tag()
        self.last_mean_return = ret_sum / (G_loc * E)
        return n_valid

    def _resync_after_checkpoint_load(self):
        """policy.load(path) (an engine extension: the reference's GaussianActor has no load(), so its GRPO runs
        cannot resume at all) restores the weights a checkpoint holds -- the state right after a learn(), where
        old_policy had just been synchronised (grpo.py:148).  The resumed run continues like the uninterrupted one
        only if old_policy is restored with them.  Plain load_state_dict() calls keep the reference's semantics
        (old_policy untouched)."""
        loads = getattr(self.policy, "checkpoint_loads", 0)
        if loads != self._seen_loads:
            self._seen_loads = loads
            self.old_policy.load_state_dict(self.policy.state_dict())
            self._synced_tag = self.policy.param_tag()
            self._old_tag = self.old_policy.param_tag()

    def save(self, path: str) -> None:
        if L.is_checkpoint_writer():
            torch.save(self.optimizer.state_dict(), os.path.join(path, "optimizer.pth"))

    def load(self, path: str) -> None:
        """grpo.py:156-160."""
        self.optimizer.load_state_dict(torch.load(os.path.join(path, "optimizer.pth"), weights_only=True))

    def metadata(self):
        return {"algorithm": "GRPO", "epsilon": self.epsilon, "beta": self.beta,
                "updates_per_iter": self.updates_per_iter}


class PPO(Algorithm):
    """algorithms/ppo.py:8-225: full-batch updates (batch_size=None, the shipped pipelines' setting; shards over
    GPUs) and randperm minibatches (batch_size=k, single GPU)."""

    def __init__(self, epsilon: float, policy, optimizer, ref_model, updates_per_iter: int, c1: float = 0.5,
                 kl_coeff: float = 0.5, gamma: float = 0.99, lam: float = 0.95, entropy: float = 0.01,
                 batch_size: int = 64, monte_carlo: bool = True):
        self.epsilon, self.c1, self.policy, self.ref_model = epsilon, c1, policy, ref_model
        self.updates_per_iter, self.optimizer, self.gamma, self.lam = updates_per_iter, optimizer, gamma, lam
        self.entropy, self.batch_size, self.kl_coeff, self.monte_carlo = entropy, batch_size, kl_coeff, monte_carlo
        if not hasattr(policy, "critic"):
            raise L.EngineError("PPO needs a GaussianActorCritic_NeuralNetwork policy")
        self.old_policy = copy.deepcopy(self.policy)                 # ppo.py:62
        self._flat_opt = _FlatOptimizer(optimizer, policy)
        self.last_stats = None

    def learn(self, buffer) -> None:
        world = _dist_world()
        if self.batch_size is not None and world > 1:
            # ppo.py:149 draws ONE torch.randperm over all valid steps; it has no sharded equivalent
            raise L.EngineError("minibatched PPO (batch_size != None) runs on one GPU; use batch_size=None when sharded")
        r = _rollout_of(buffer)
        pol = self.policy
        flat = pol.flat_parameters()
        na = pol.actor.n_params()
        a_flat, c_flat = flat[:na], flat[na:]
        a_dims, c_dims, act_name, cov = pol.actor.dims, pol.critic.dims, pol.actor.activation_name, pol.cov_diag
        # ppo.py:93-94: critic values over the rollout
        values, _ = engine.policy_forward_traj(c_dims, act_name, c_flat, r.obs, None, None, r.len, want_mu=True,
                                               want_logp=False)
        values = values.view(r.T, r.N)
        mode = L.ADV_PPO_MC if self.monte_carlo else L.ADV_PPO_GAE
        # ppo.py:100-139.  The z-scores are over every valid step of the WHOLE rollout: a rank that holds
        # a shard (whole groups per GPU) computes its raw values and five additive sums, the sums are
        # allreduced (40 bytes), and every rank normalises with the global statistics.
        adv, rtg, sums = engine.advantage_ppo_raw(mode, r.G, r.E, r.T, self.gamma, self.lam, r.rew, r.len, values)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(sums)
        engine.advantage_ppo_normalize(r.T, r.len, sums, adv, rtg)
        # ppo.py:142-143: old log-prob from self.policy at the start of learn()
        _, old_logp = engine.policy_forward_traj(a_dims, act_name, a_flat, r.obs, cov, r.act, r.len)
        n_valid = int(round(float(sums[4].item())))                 # global valid-step count (the .mean()s)
        grad = torch.empty_like(flat)
        comm = self._flat_opt.peer_comm(flat) if (world > 1 and self.batch_size is None) else None
        if self.batch_size is not None:
            self._learn_minibatched(r, flat, grad, na, a_dims, c_dims, act_name, cov, adv, rtg, old_logp, n_valid)
            self.old_policy.load_state_dict(self.policy.state_dict())   # ppo.py:186
            return
        with engine.length_order(r.len, r.T):
            for _ in range(self.updates_per_iter):                      # ppo.py:147 (full batch; the order of a
                # permutation does not change a mean)
                if comm is not None:
                    grad = comm.grad_slot()[:flat.numel()]               # this rank's slot of the peer-memory window
                _, stats = engine.policy_grad(a_dims, act_name, a_flat, cov, r.obs, r.act, adv, old_logp, r.len,
                                              self.epsilon, -1.0 / n_valid, self.kl_coeff / n_valid,
                                              out_grad=grad[:na])        # :160-166, 175-176
                engine.value_grad(c_dims, act_name, c_flat, r.obs, rtg, r.len, self.c1 / n_valid, out_grad=grad[na:])
                if comm is not None:
                    self._flat_opt.step_allreduced(flat, comm)          # actor + critic gradients, allreduce + Adam fused
                else:
                    if world > 1:
                        import torch.distributed as dist
                        dist.all_reduce(grad)                            # actor + critic gradients in one message
                    self._flat_opt.step(flat, grad)                     # :181-183 (entropy term has zero gradient)
                pol.bump_param_epoch()
                self.last_stats = stats
        self.old_policy.load_state_dict(self.policy.state_dict())   # ppo.py:186

    def _learn_minibatched(self, r, flat, grad, na, a_dims, c_dims, act_name, cov, adv, rtg, old_logp, n_valid):
        """ppo.py:147-183 with batch_size != None: one `torch.randperm(data_size)` per epoch from torch's
        default CPU generator (the reference's call, so a seeded run draws the same permutation), sliced
        into minibatches over the valid steps in the reference's flattening order (env-major, step-minor);
        one Adam step per minibatch.  Each minibatch is a list of flat slot ids for the *_batch kernels."""
        a_flat, c_flat = flat[:na], flat[na:]
        dev = flat.device
        lens = r.len.to(torch.int64)
        csum = torch.cumsum(lens, 0)
        row0 = csum - lens
        for _ in range(self.updates_per_iter):
            permutation = torch.randperm(n_valid)                   # ppo.py:149
            perm_d = permutation.to(dev)
            n_of = torch.searchsorted(csum, perm_d, right=True)     # valid-step index k -> env n, step t
            sid_all = ((perm_d - row0[n_of]) * r.N + n_of).contiguous()
            for start in range(0, n_valid, self.batch_size):        # ppo.py:152-153
                sid = sid_all[start:start + self.batch_size]
                m = sid.numel()
                _, stats = engine.policy_grad_batch(a_dims, act_name, a_flat, cov, r.obs, r.act, adv, old_logp, sid,
                                                    self.epsilon, -1.0 / m, self.kl_coeff / m, out_grad=grad[:na])
                engine.value_grad_batch(c_dims, act_name, c_flat, r.obs, rtg, sid, self.c1 / m, out_grad=grad[na:])
                self._flat_opt.step(flat, grad)
                self.policy.bump_param_epoch()
                self.last_stats = stats

    def metadata(self) -> dict:
        return {"algorithm": "PPO", "epsilon": self.epsilon, "c1": self.c1, "kl_coeff": self.kl_coeff,
                "gamma": self.gamma, "lam": self.lam, "entropy": self.entropy, "batch_size": self.batch_size,
                "updates_per_iter": self.updates_per_iter}

    def save(self, path: str) -> None:
        if L.is_checkpoint_writer():
            torch.save(self.optimizer.state_dict(), os.path.join(path, "optimizer.pt"))

    def load(self, path: str) -> None:
        """ppo.py:217-225."""
        self.optimizer.load_state_dict(torch.load(os.path.join(path, "optimizer.pt"), weights_only=True))
        self.old_policy.load_state_dict(self.policy.state_dict())
