"""CPU port of the reference's hot loop, for TIMING the CPU way (bench.py's
`cpu_baseline` leg and `--impl reference`).  TEST/BENCH INFRASTRUCTURE ONLY.

The reference itself (pure Python, /root/reference) cannot travel to the GPU
box, so this module restates its execution STRUCTURE, not just its arithmetic:

  * rollout: one OS process per worker (rollout/rollout_manager.py:44-57), each
    running E sequential episodes (rollout/rollout_worker.py:43-75); every step
    is one single-row torch MLP forward plus a freshly built
    `torch.distributions.MultivariateNormal` (policies/actor_critic.py:124-136)
    followed by a float64 numpy env step (oracle/restate.py, batch of 1) --
    with OMP_NUM_THREADS=1 per worker, as BASELINE.md section 3 prescribes;
  * update: `GRPO.learn`'s loop over groups with boolean-mask gathers, two
    forwards and one backward per group on torch-CPU autograd, then Adam
    (algorithms/grpo.py:106-148).
"""
from __future__ import annotations

import math
import os
import time

import numpy as np

import restate as R


def _make_net(Ws, bs):
    import torch
    layers = []
    for i, (W, b) in enumerate(zip(Ws, bs)):
        lin = torch.nn.Linear(W.shape[1], W.shape[0])
        with torch.no_grad():
            lin.weight.copy_(torch.from_numpy(np.asarray(W)))
            lin.bias.copy_(torch.from_numpy(np.asarray(b)))
        layers.append(lin)
        if i < len(Ws) - 1:
            layers.append(torch.nn.ReLU())
    return torch.nn.Sequential(*layers)


def _worker(args):
    """One RolloutWorker.run_episodes (rollout_worker.py:19-84) in its own process."""
    import torch
    from torch.distributions import MultivariateNormal
    torch.set_num_threads(1)
    kind, T, dt, Ws, bs, cov, E, seed, restart = args
    cfg = R.EnvCfg(kind, T, dt)
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    net = _make_net(Ws, bs)
    covm = torch.diag(torch.tensor([float(c) for c in cov]))
    O, A = R.OBS_DIM[kind], R.ACT_DIM[kind]
    obs = np.zeros((E, T, O)); act = np.zeros((E, T, A)); rew = np.zeros((E, T)); mask = np.zeros((E, T))
    lens = np.zeros(E, int)
    s_init = R.reset_states(kind, 1, rng)
    for e in range(E):
        s = s_init.copy() if restart else R.reset_states(kind, 1, rng)
        steps = np.zeros(1, np.int64); bal = np.zeros(1, np.int64)
        done, t = False, 0
        while not done and t < T:
            obs[e, t] = s[0]
            x = torch.from_numpy(s[0]).float()                        # actor_critic.py:124-125
            mean = net(x)                                             # no no_grad, as the reference
            dist = MultivariateNormal(mean, covm)                     # rebuilt every step (:131)
            a = dist.sample()
            dist.log_prob(a)                                          # computed and discarded (:136)
            a_np = a.detach().numpy()
            s, r, d, bal = R.env_step(cfg, s, a_np[None, :], steps, bal, np.float64)
            steps = steps + 1
            act[e, t] = a_np; rew[e, t] = r[0]
            done = bool(d[0]); t += 1
        lens[e] = t; mask[e, :t] = 1
    return obs.astype(np.float32), act.astype(np.float32), rew.astype(np.float32), lens, mask.astype(np.float32)


def rollout_mp(kind, T, dt, Ws, bs, cov, G, E, restart, seed, procs=None):
    """G workers over `procs` processes; returns ([G,E,T,.] arrays, seconds, procs used)."""
    import multiprocessing as mp
    procs = procs or min(G, os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    jobs = [(kind, T, dt, Ws, bs, cov, E, seed + g, restart) for g in range(G)]
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        pool.map(_noop, range(procs))                                 # start the workers before timing
        t0 = time.perf_counter()
        res = pool.map(_worker, jobs, chunksize=1)
        dt_s = time.perf_counter() - t0
    out = [np.stack([r[i] for r in res]) for i in range(5)]
    return out, dt_s, procs


def _noop(_):
    return 0


def grpo_learn(obs, act, rew, mask, Ws, bs, cov, gamma, eps_clip, updates, lr):
    """GRPO.learn (grpo.py:50-148) on torch-CPU: RTG loop over T, Python loop over
    groups, autograd backward, torch.optim.Adam.  Returns seconds."""
    import torch
    t0 = time.perf_counter()
    net = _make_net(Ws, bs)
    old = _make_net(Ws, bs)
    opt = torch.optim.Adam(net.parameters(), lr=lr)
    o, a, r, m = (torch.from_numpy(x) for x in (obs, act, rew, mask))
    G, E, T, _ = o.shape
    rtg = torch.zeros_like(r)
    for i in reversed(range(T)):                                      # grpo.py:66-74
        if i < T - 1:
            rtg[:, :, i] = r[:, :, i] * m[:, :, i] + gamma * rtg[:, :, i + 1] * m[:, :, i + 1]
        else:
            rtg[:, :, i] = r[:, :, i] * m[:, :, i]
    o, a, rtg, m = o.reshape(G, E * T, -1), a.reshape(G, E * T, -1), rtg.reshape(G, -1), m.reshape(G, -1)
    cd = torch.tensor([float(c) for c in cov])
    sd = torch.sqrt(cd)
    A = a.shape[-1]

    def logp(n, oo, aa):
        z = (aa - n(oo)) / sd
        return -0.5 * (A * math.log(2 * math.pi) + (z * z).sum(-1)) - torch.log(sd).sum()

    for _ in range(updates):
        J = 0
        for g in range(G):                                            # grpo.py:108-137
            sel = m[g].bool()
            og, ag, rg = o[g][sel], a[g][sel], rtg[g][sel]
            Ag = (rg - rg.mean()) / torch.std(rg + 1e-8)
            with torch.no_grad():
                olp = logp(old, og, ag)
            lp = logp(net, og, ag)
            ratio = torch.exp(lp - olp)
            J = J + torch.min(ratio * Ag, torch.clamp(ratio, 1 - eps_clip, 1 + eps_clip) * Ag).sum()
        J = J / G
        opt.zero_grad(); J.backward(); opt.step()
    return time.perf_counter() - t0
