"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

TEST INFRASTRUCTURE ONLY -- runs in the build container, where
/root/reference is mounted; the fixtures it writes are committed and are what
the GPU box (which has no /root/reference) checks against.

    python oracle/make_golden.py            # (re)writes tests/golden/*.npz

What is driven (no edits to the reference; injection points per SURVEY.md
Appendix A):
  * env transitions: the real ``Env.step`` of CartPole / Pendulum / QuadPole2D /
    QuadPole on prescribed float32 raw-action sequences, plus
    ``Quadrotor._dynamics`` known-answer vectors;
  * policy: ``GaussianActor_NeuralNetwork.forward/log_prob`` with the standard
    normal draw replaced by a supplied tensor;
  * rollouts: ``RolloutManager(use_multiprocessing=False)`` with injected
    initial states (``reset`` wrapped) and injected noise;
  * updates: ``GRPO.learn`` / ``PPO.learn`` with ``SGD(lr=1)`` (gradient =
    theta_before - theta_after) and with ``Adam``.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _envs():
    from environments.cartpole_env import CartPole
    from environments.pendulum_env import Pendulum
    from environments.quadrotor_env import QuadPole, QuadPole2D
    return {0: CartPole, 1: Pendulum, 2: QuadPole2D, 3: QuadPole}


STATE_KEYS = {0: ["cartpole"], 1: ["pendulum"], 2: ["quadrotor", "pendulum"], 3: ["quadrotor", "pendulum"]}
SPLIT = {0: [5], 1: [3], 2: [7, 3], 3: [13, 7]}


def set_state(env, kind, vec):
    """Put a full observation vector into env.state_dict (the layout reset() builds)."""
    off = 0
    for key, n in zip(STATE_KEYS[kind], SPLIT[kind]):
        env.state_dict[key] = np.array(vec[off:off + n], dtype=np.float64)
        off += n
    env._initial_state = {k: v.copy() for k, v in env.state_dict.items()}
    env._steps = 0
    env._time = 0
    env._time_balanced = 0


def gen_transitions(kind, rng, episodes, scale, max_steps=None, pd=False, env_kw=None, dynamics=False):
    """Run episodes of the real env on random raw actions; record every transition.  `dynamics`: also call
    Env._dynamics(state, wrapped action) directly (the known answers of the tg_env_dynamics entry point)."""
    cls = _envs()[kind]
    kw = dict(env_kw or {})
    if max_steps is not None:
        kw["max_steps"] = max_steps
    env = cls(**kw)
    rows = {k: [] for k in ["state", "action", "next", "reward", "done", "steps_done", "bal_count"]}
    if dynamics:
        rows["control"], rows["dyn_next"] = [], []
    for ep in range(episodes):
        obs, _ = env.reset()
        bal = 0
        for t in range(env.max_steps):
            A = env.action_space.shape[0]
            if pd and kind == 1:
                th = np.arctan2(obs[0], obs[1])
                delta = np.arctan2(np.sin(th - np.pi), np.cos(th - np.pi))
                a = np.array([-8.0 * delta - 2.0 * obs[2]], dtype=np.float32)
            else:
                a = (scale * rng.standard_normal(A)).astype(np.float32)
            rows["state"].append(np.array(obs, dtype=np.float64))
            rows["action"].append(a)
            rows["steps_done"].append(t)
            rows["bal_count"].append(bal)
            if dynamics:
                u = np.asarray(env._wrap_action(a), dtype=np.float32)
                rows["control"].append(u)
                rows["dyn_next"].append(np.asarray(env._dynamics(np.array(obs, dtype=np.float64), u), dtype=np.float64).reshape(-1))
            obs, r, f1, f2, _ = env.step(a)
            if kind == 1:
                bal = bal + 1 if obs[1] <= -0.99 else 0
                assert (bal > 0) == (env._time_balanced > 0)
            rows["next"].append(np.array(obs, dtype=np.float64))
            rows["reward"].append(float(r))
            rows["done"].append(bool(f1 or f2))
            if f1 or f2:
                break
    return {k: np.array(v) for k, v in rows.items()}


def gen_quadrotor12(rng):
    from environments.quadrotor_env import Quadrotor
    q = Quadrotor()
    S = rng.uniform(-0.5, 0.5, (64, 12))
    U = rng.uniform(2.0, 3.0, (64, 4))
    out = np.stack([np.asarray(q._dynamics(S[i], U[i])) for i in range(64)])
    return {"state": S, "control": U, "next": out}


class Injector:
    """Feeds injected initial states and noise to the real worker loop."""

    def __init__(self, kind, init_states, noise, E, restart):
        self.kind, self.init, self.noise, self.E, self.restart = kind, init_states, noise, E, restart
        self.cur = None  # (worker, episode, t)

    def wrap_env(self, env, worker):
        inj = self
        orig_step, orig_restart = env.step, env.restart
        st = {"ep": -1}

        def reset():
            st["ep"] += 1
            ep = min(st["ep"], inj.E - 1)
            n = worker * inj.E + (0 if inj.restart else ep)
            set_state(env, inj.kind, inj.init[n])
            inj.cur = [worker, ep, 0]
            return env._get_obs(), env._get_info()

        def restart():
            st["ep"] += 1
            inj.cur = [worker, min(st["ep"], inj.E - 1), 0]
            return orig_restart()

        def step(a):
            out = orig_step(a)
            inj.cur[2] += 1
            return out

        env.reset, env.restart, env.step = reset, restart, step
        return env

    def standard_normal(self, shape, dtype, device):
        w, e, t = self.cur
        return torch.from_numpy(self.noise[t, w * self.E + e].copy()).to(dtype).reshape(shape)


def make_policy(kind, hidden, cov, seed, critic=False, activation="ReLU", weights=None):
    """`weights`: path (relative to the reference root) of a shipped policy.pt to load (reports/**)."""
    from policies.actor_critic import GaussianActor_NeuralNetwork, GaussianActorCritic_NeuralNetwork
    O = {0: 5, 1: 3, 2: 10, 3: 20}[kind]
    A = {0: 1, 1: 1, 2: 2, 3: 4}[kind]
    torch.manual_seed(seed)
    cls = GaussianActorCritic_NeuralNetwork if critic else GaussianActor_NeuralNetwork
    pol = cls(O, A, hidden, activation, cov)
    if weights is not None:
        sd = torch.load(os.path.join(ref_shims.REFERENCE_ROOT, weights), weights_only=True)
        if critic:
            pol.actor.load_state_dict(sd["actor"]); pol.critic.load_state_dict(sd["critic"])
        else:
            pol.actor.load_state_dict(sd)
    return pol


def policy_arrays(policy):
    sd = policy.actor.state_dict()
    keys = sorted({int(k.split(".")[1]) for k in sd})
    out = {}
    for i, k in enumerate(keys):
        out[f"W{i}"] = sd[f"network.{k}.weight"].numpy().copy()
        out[f"b{i}"] = sd[f"network.{k}.bias"].numpy().copy()
    if hasattr(policy, "critic"):
        sd = policy.critic.state_dict()
        for i, k in enumerate(keys):
            out[f"cW{i}"] = sd[f"network.{k}.weight"].numpy().copy()
            out[f"cb{i}"] = sd[f"network.{k}.bias"].numpy().copy()
    return out


def run_rollout(kind, policy, init, noise, G, E, T, restart):
    import torch.distributions.multivariate_normal as mvn
    from rollout.rollout_manager import RolloutManager
    cls = _envs()[kind]
    inj = Injector(kind, init, noise, E, restart)
    counter = {"i": -1}

    def env_fn():
        counter["i"] += 1
        env = cls(max_steps=T)
        if counter["i"] == 0:
            return env  # the manager's own shape-discovery env
        return inj.wrap_env(env, counter["i"] - 1)

    saved = mvn._standard_normal
    mvn._standard_normal = inj.standard_normal
    try:
        mgr = RolloutManager(env_fn, policy, restart=restart, num_workers=G,
                             num_episodes_per_worker=E, use_multiprocessing=False)
        obs, act, rew, ln, mask = mgr.rollout()
    finally:
        mvn._standard_normal = saved
    return obs.numpy(), act.numpy(), rew.numpy(), ln.numpy(), mask.numpy()


class Buf:
    pass


KINK_MARGIN = 2e-6


def kink_margin_over_updates(out, act_ids, lr=3e-4, updates=6):
    """min ReLU-kink margin (restate.kink_margin) over the valid samples of a fixture, at its initial weights and
    after each of the Adam updates the fixture records (replayed with the oracle)."""
    import restate as R
    Ws, bs, i = [], [], 0
    while f"W{i}" in out:
        Ws.append(out[f"W{i}"]); bs.append(out[f"b{i}"]); i += 1
    cov = np.full(out["act"].shape[-1], out["cov"], np.float32)
    x = out["obs"][out["mask"] > 0]
    _, adv = R.grpo_advantage(out["rew"], out["mask"], float(out["gamma"]))
    params = []
    for w, b in zip(Ws, bs):
        params += [w.copy(), b.copy()]
    m, v = [np.zeros_like(p) for p in params], [np.zeros_like(p) for p in params]
    old = [p.copy() for p in params]
    margin0 = margin = R.kink_margin(x, params[0::2], params[1::2], act_ids)
    if margin0 < KINK_MARGIN:
        return margin0, margin0
    for step in range(1, updates + 1):
        _, dW, db, _, _ = R.grpo_objective_and_grad(out["obs"], out["act"], adv, out["mask"], params[0::2], params[1::2],
                                                    old[0::2], old[1::2], cov, float(out["eps_clip"]), act=act_ids,
                                                    dtype="float32")
        grads = []
        for a, b in zip(dW, db):
            grads += [a, b]
        params = R.adam_step(params, grads, m, v, step, lr)
        if step % 3 == 0:
            old = [p.copy() for p in params]
        margin = min(margin, R.kink_margin(x, params[0::2], params[1::2], act_ids))
    return margin0, margin


def gen_rollout_and_learn(screen=True, **kw):
    """_gen_rollout_and_learn with the numpy seed screened: a fixture whose samples come within KINK_MARGIN of a
    ReLU kink is regenerated with seed + 1000 -- there the fp32 answer depends on the summation order
    (tie-breaking), which is not what these fixtures pin.  Screened: the initial weights always (gradient tests);
    the weights after each recorded Adam update too for widths <= 64 (at 128 / 256 the ~3 million pre-activations
    of the six updates cannot all clear the margin; `kink_margin` is recorded and the Adam tests use the
    kink-tolerant criterion when it is small)."""
    import restate as R
    act = kw.get("activation", "ReLU")
    act_ids = R.acts_from_names(act if isinstance(act, str) else ",".join(act))
    wide = max(kw["hidden"], default=0) > 64
    for attempt in range(200):
        out = _gen_rollout_and_learn(**kw)
        m0, margin = kink_margin_over_updates(out, act_ids)
        out["kink_margin0"], out["kink_margin"] = np.float64(m0), np.float64(margin)
        if not screen or (m0 >= KINK_MARGIN and (wide or margin >= KINK_MARGIN)):
            out["seed_used"] = np.int64(kw["seed"])
            return out
        print(f"    seed {kw['seed']}: kink margin {m0:.1e} / {margin:.1e} < {KINK_MARGIN:.0e}, trying seed {kw['seed'] + 1000}")
        kw = dict(kw, seed=kw["seed"] + 1000)
    raise RuntimeError("no seed with a clear kink margin")


def _gen_rollout_and_learn(kind, hidden, cov, G, E, T, restart, seed, gamma, eps_clip, noise_scale=1.0,
                           activation="ReLU", weights=None):
    rng = np.random.default_rng(seed)
    sys.path.insert(0, HERE)
    import restate
    policy = make_policy(kind, hidden, cov, seed, activation=activation, weights=weights)
    A = {0: 1, 1: 1, 2: 2, 3: 4}[kind]
    n_init = G if restart else G * E
    init_small = restate.reset_states(kind, n_init, rng)
    init = np.repeat(init_small, E, axis=0) if restart else init_small
    noise = (noise_scale * rng.standard_normal((T, G * E, A))).astype(np.float32)
    obs, act, rew, ln, mask = run_rollout(kind, policy, init, noise, G, E, T, restart)
    out = dict(kind=kind, hidden=np.array(hidden, dtype=np.int64), cov=np.float32(cov), G=G, E=E, T=T,
               restart=restart, gamma=gamma, eps_clip=eps_clip, init=init, noise=noise,
               obs=obs, act=act, rew=rew, len=ln, mask=mask, **policy_arrays(policy))
    if activation != "ReLU":
        out["activation"] = np.array(activation if isinstance(activation, str) else ",".join(activation))

    # --- log_prob KAT on the rollout's own samples
    with torch.no_grad():
        sel = torch.from_numpy(mask).bool()
        lp, ent = policy.log_prob(torch.from_numpy(obs)[sel], torch.from_numpy(act)[sel])
    out["logp_valid"] = lp.numpy()
    out["entropy"] = float(ent.reshape(-1)[0])

    # --- GRPO.learn: gradient through SGD(lr=1), one update
    from algorithms.grpo import GRPO
    buf = Buf()
    buf.group_observations, buf.group_actions = torch.from_numpy(obs), torch.from_numpy(act)
    buf.group_rewards, buf.group_masks = torch.from_numpy(rew), torch.from_numpy(mask)
    import copy
    p1 = copy.deepcopy(policy)
    before = [p.detach().clone() for p in p1.parameters()]
    algo = GRPO(eps_clip, 0.0, gamma, p1, torch.optim.SGD(p1.parameters(), lr=1.0), None, updates_per_iter=1)
    algo.learn(buf)
    for i, (b, a) in enumerate(zip(before, p1.parameters())):
        out[f"grpo_grad{i}"] = (b - a.detach()).numpy()

    # --- GRPO.learn with Adam, 3 updates (exercises ratio != 1 and the clip)
    p2 = copy.deepcopy(policy)
    algo = GRPO(eps_clip, 0.0, gamma, p2, torch.optim.Adam(p2.parameters(), lr=3e-4), None, updates_per_iter=3)
    algo.learn(buf)
    for i, a in enumerate(p2.parameters()):
        out[f"grpo_adam3_p{i}"] = a.detach().numpy().copy()
    # a second learn() call on the same buffer: old_policy has been synced (:148)
    algo.learn(buf)
    for i, a in enumerate(p2.parameters()):
        out[f"grpo_adam6_p{i}"] = a.detach().numpy().copy()
    return out


def gen_ppo(kind, hidden, cov, G, E, T, seed, gamma, lam, eps_clip, monte_carlo, weights=None):
    """PPO.learn full-batch (batch_size=None), SGD(lr=1) one update and Adam 3 updates."""
    import copy
    from algorithms.ppo import PPO
    rng = np.random.default_rng(seed)
    import restate
    policy = make_policy(kind, hidden, cov, seed, critic=True, weights=weights)
    A = {0: 1, 1: 1, 2: 2, 3: 4}[kind]
    init = restate.reset_states(kind, G * E, rng)
    noise = rng.standard_normal((T, G * E, A)).astype(np.float32)
    obs, act, rew, ln, mask = run_rollout(kind, policy, init, noise, G, E, T, False)
    out = dict(kind=kind, hidden=np.array(hidden), cov=np.float32(cov), G=G, E=E, T=T, gamma=gamma,
               lam=lam, eps_clip=eps_clip, monte_carlo=monte_carlo, init=init, noise=noise,
               obs=obs, act=act, rew=rew, len=ln, mask=mask, **policy_arrays(policy))
    buf = Buf()
    buf.group_observations, buf.group_actions = torch.from_numpy(obs), torch.from_numpy(act)
    buf.group_rewards, buf.group_masks = torch.from_numpy(rew), torch.from_numpy(mask)
    kw = dict(c1=0.5, kl_coeff=0.5, gamma=gamma, lam=lam, entropy=0.01, batch_size=None, monte_carlo=monte_carlo)
    p1 = copy.deepcopy(policy)
    before = [p.detach().clone() for p in p1.parameters()]
    PPO(eps_clip, p1, torch.optim.SGD(p1.parameters(), lr=1.0), None, 1, **kw).learn(buf)
    for i, (b, a) in enumerate(zip(before, p1.parameters())):
        out[f"ppo_grad{i}"] = (b - a.detach()).numpy()
    p2 = copy.deepcopy(policy)
    PPO(eps_clip, p2, torch.optim.Adam(p2.parameters(), lr=2e-4), None, 3, **kw).learn(buf)
    for i, a in enumerate(p2.parameters()):
        out[f"ppo_adam3_p{i}"] = a.detach().numpy().copy()
    return out


def gen_ppo_minibatch(kind, hidden, cov, G, E, T, seed, gamma, lam, eps_clip, batch_size, updates, torch_seed):
    """PPO.learn with randperm minibatches (ppo.py:147-183): Adam, `updates` epochs; torch's default CPU
    generator is seeded right before learn() so the permutations can be redrawn by the checker."""
    import copy
    from algorithms.ppo import PPO
    rng = np.random.default_rng(seed)
    import restate
    policy = make_policy(kind, hidden, cov, seed, critic=True)
    A = {0: 1, 1: 1, 2: 2, 3: 4}[kind]
    init = restate.reset_states(kind, G * E, rng)
    noise = rng.standard_normal((T, G * E, A)).astype(np.float32)
    obs, act, rew, ln, mask = run_rollout(kind, policy, init, noise, G, E, T, False)
    out = dict(kind=kind, hidden=np.array(hidden), cov=np.float32(cov), G=G, E=E, T=T, gamma=gamma,
               lam=lam, eps_clip=eps_clip, monte_carlo=False, init=init, noise=noise, batch_size=batch_size,
               updates=updates, torch_seed=torch_seed,
               obs=obs, act=act, rew=rew, len=ln, mask=mask, **policy_arrays(policy))
    buf = Buf()
    buf.group_observations, buf.group_actions = torch.from_numpy(obs), torch.from_numpy(act)
    buf.group_rewards, buf.group_masks = torch.from_numpy(rew), torch.from_numpy(mask)
    kw = dict(c1=0.5, kl_coeff=0.5, gamma=gamma, lam=lam, entropy=0.01, batch_size=batch_size, monte_carlo=False)
    p2 = copy.deepcopy(policy)
    algo = PPO(eps_clip, p2, torch.optim.Adam(p2.parameters(), lr=2e-4), None, updates, **kw)
    torch.manual_seed(torch_seed)
    algo.learn(buf)
    for i, a in enumerate(p2.parameters()):
        out[f"ppo_mb_adam_p{i}"] = a.detach().numpy().copy()
    return out


SHIPPED = ["CartPole/cartpole_nn_grpo/001", "CartPole/cartpole_nn_ppo/001", "QuadPole2D/quadpole2d_nn_ppo/001"]


def copy_shipped():
    """The reference ships three trained checkpoints under reports/** (policy.pt, optimizer.pt[h], reward.csv,
    metadata.json: DATA files written by Pipeline.save, pipelines/pipeline.py:104-118).  They are copied to
    tests/golden/shipped/ so that the GPU box (no /root/reference) can test that they load and resume."""
    import shutil
    for rel in SHIPPED:
        src = os.path.join(ref_shims.REFERENCE_ROOT, "reports", rel)
        dst = os.path.join(OUT, "shipped", rel)
        os.makedirs(dst, exist_ok=True)
        for f in os.listdir(src):
            if f.split(".")[-1] in ("pt", "pth", "csv", "json"):
                shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))


def gen_resume(algo_name, kind, hidden, report, G, E, T, seed, updates):
    """Resume from a shipped checkpoint the way Pipeline.load does (pipelines/pipeline.py:93-102:
    algorithm.load, policy.load, buffer.load), then one learn() on an injected rollout: the loaded Adam
    moments and step count decide the update, so the weights afterwards pin the whole resume path.
    (GaussianActor_NeuralNetwork has no load() in the reference -- GRPO resume raises there; its policy.pt is
    loaded through load_state_dict instead.)"""
    import json
    from algorithms.grpo import GRPO
    from algorithms.ppo import PPO
    from buffers.rollout_buffer import Rollout_Buffer
    path = os.path.join(ref_shims.REFERENCE_ROOT, "reports", report)
    meta = json.load(open(os.path.join(path, "metadata.json")))
    rng = np.random.default_rng(seed)
    import restate
    ppo = algo_name == "ppo"
    policy = make_policy(kind, hidden, 0.5, seed, critic=ppo)
    am = meta["algorithm"]
    if ppo:
        opt = torch.optim.Adam(policy.parameters(), lr=2e-4)
        algo = PPO(am["epsilon"], policy, opt, None, updates, c1=am["c1"], kl_coeff=am["kl_coeff"], gamma=am["gamma"],
                   lam=am["lam"], entropy=am["entropy"], batch_size=am["batch_size"], monte_carlo=True)
        algo.load(path)
        policy.load(path)
    else:
        opt = torch.optim.Adam(policy.parameters(), lr=3e-4)
        algo = GRPO(am["epsilon"], am["beta"], 0.5, policy, opt, None, updates_per_iter=updates)
        algo.load(path)
        policy.load_state_dict(torch.load(os.path.join(path, "policy.pt"), weights_only=True))
        algo.old_policy.load_state_dict(policy.state_dict())
    rb = Rollout_Buffer.__new__(Rollout_Buffer)
    n_epochs = Rollout_Buffer.load(rb, path)
    A = {0: 1, 1: 1, 2: 2, 3: 4}[kind]
    init = restate.reset_states(kind, G * E, rng)
    noise = rng.standard_normal((T, G * E, A)).astype(np.float32)
    obs, act, rew, ln, mask = run_rollout(kind, policy, init, noise, G, E, T, False)
    buf = Buf()
    buf.group_observations, buf.group_actions = torch.from_numpy(obs), torch.from_numpy(act)
    buf.group_rewards, buf.group_masks = torch.from_numpy(rew), torch.from_numpy(mask)
    step0 = int(float(opt.state_dict()["state"][0]["step"]))
    algo.learn(buf)
    out = dict(kind=kind, hidden=np.array(hidden, dtype=np.int64), G=G, E=E, T=T, updates=updates, report=np.array(report),
               init=init, noise=noise, obs=obs, act=act, rew=rew, len=ln, mask=mask, n_epochs_loaded=n_epochs,
               step_loaded=step0, step_after=int(float(opt.state_dict()["state"][0]["step"])),
               gamma=float(am.get("gamma", 0.5)), lam=float(am.get("lam", 0.0)), epsilon=float(am["epsilon"]))
    for i, a in enumerate(policy.parameters()):
        out[f"resume_p{i}"] = a.detach().numpy().copy()
    return out


def _same(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b, equal_nan=a.dtype.kind == "f")


def emit(name, arrays, check, report):
    """Write tests/golden/<name>.npz, or (--check) compare with the committed file bit for bit."""
    path = os.path.join(OUT, name + ".npz")
    if check:
        old = dict(np.load(path, allow_pickle=False))
        new = {k: np.asarray(v) for k, v in arrays.items()}
        bad = sorted(set(old) ^ set(new)) + [k for k in old if k in new and not _same(old[k], new[k])]
        report[name] = bad
        print(f"{name}: {'bit-identical' if not bad else 'DIFFERS in ' + str(bad)}")
    else:
        np.savez_compressed(path, **arrays)


# name -> (generator, kwargs).  Every generator seeds its own numpy Generator / torch seed, and main()
# seeds numpy's GLOBAL RNG per fixture (the reference's env.reset() draws from it), so each file
# regenerates bit-identically on its own (`--only name`, `--check`).
CP_GRPO = "reports/CartPole/cartpole_nn_grpo/001/policy.pt"
CP_PPO = "reports/CartPole/cartpole_nn_ppo/001/policy.pt"
QP2_PPO = "reports/QuadPole2D/quadpole2d_nn_ppo/001/policy.pt"


def fixtures():
    F = {}

    def transitions(kind):
        def gen():
            rng = np.random.default_rng(20261018 + kind)
            if kind == 1:
                a = gen_transitions(1, rng, episodes=2, scale=0.8, max_steps=200)
                b = gen_transitions(1, rng, episodes=1, scale=0.0, max_steps=200, pd=True)   # 101 balanced steps
                return {k: np.concatenate([a[k], b[k]]) for k in a}
            return gen_transitions(kind, rng, episodes=6, scale=1.0, max_steps={0: 120, 2: 150, 3: 150}[kind])
        return gen
    for k in range(4):
        F[f"transitions_env{k}"] = transitions(k)
    F["quadrotor12_dynamics"] = lambda: gen_quadrotor12(np.random.default_rng(20261019))
    # non-default physical constructor arguments (cartpole_env.py:7-16, pendulum_env.py:8-17) and direct
    # Env._dynamics(state, control) calls for all four envs
    CP_KW = dict(masscart=0.7, masspole=0.3, length=0.8, gravity=9.5, timestep=0.025)
    PD_KW = dict(mass=0.6, length=0.9, gravity=9.7, timestep=0.04)
    F["transitions_cartpole_params"] = lambda: dict(gen_transitions(0, np.random.default_rng(41), 4, 1.0, 100, env_kw=CP_KW,
                                                                    dynamics=True), **{k: np.float64(v) for k, v in CP_KW.items()})
    F["transitions_pendulum_params"] = lambda: dict(gen_transitions(1, np.random.default_rng(42), 3, 0.8, 100, env_kw=PD_KW,
                                                                    dynamics=True), **{k: np.float64(v) for k, v in PD_KW.items()})
    F["dynamics_quadpole2d"] = lambda: gen_transitions(2, np.random.default_rng(43), 2, 1.0, 60, dynamics=True)
    F["dynamics_quadpole"] = lambda: gen_transitions(3, np.random.default_rng(44), 2, 1.0, 60, dynamics=True)

    roll = {
        # round 1: small policies, one per env
        "cartpole": dict(kind=0, hidden=[32, 32], cov=0.5, G=3, E=4, T=40, restart=False, seed=1, gamma=0.5, eps_clip=0.15),
        "pendulum": dict(kind=1, hidden=[64, 64], cov=0.5, G=4, E=4, T=30, restart=True, seed=2, gamma=0.99, eps_clip=0.2),
        "quadpole2d": dict(kind=2, hidden=[32, 32], cov=0.5, G=3, E=3, T=60, restart=True, seed=3, gamma=0.99, eps_clip=0.2),
        "quadpole": dict(kind=3, hidden=[64, 64], cov=0.3, G=3, E=4, T=80, restart=True, seed=4, gamma=0.999, eps_clip=0.2),
        # round 2: the BASELINE.json shapes.  cfg 1 at FULL size with the shipped trained weights
        # (pipelines/cartpole_pipeline_grpo.py:54-76: 10 workers x 10 episodes x 500 steps, 5-128^4-1, cov 0.5,
        # eps 0.15, gamma 0.5, restart=False)
        "cartpole_cfg1": dict(kind=0, hidden=[128, 128, 128, 128], cov=0.5, G=10, E=10, T=500, restart=False, seed=21,
                              gamma=0.5, eps_clip=0.15, weights=CP_GRPO, screen=False),   # 4,370 samples x 512 units:
        # not screenable; the margin is recorded and the GPU test falls back to kink-tolerant bounds when it is small
        # cfg 3 / cfg 4 policy shapes (tensor-core kernels) at small N
        "quadpole2d_w128": dict(kind=2, hidden=[128, 128], cov=0.5, G=3, E=4, T=60, restart=True, seed=22, gamma=0.99,
                                eps_clip=0.2),
        "quadpole_w256": dict(kind=3, hidden=[256, 256], cov=0.3, G=3, E=4, T=80, restart=True, seed=23, gamma=0.999,
                              eps_clip=0.2),
        # depth 1 / 3 / 0, odd widths, other activations
        "pendulum_h1": dict(kind=1, hidden=[48], cov=0.5, G=3, E=4, T=30, restart=True, seed=24, gamma=0.99, eps_clip=0.2),
        "cartpole_h3": dict(kind=0, hidden=[32, 48, 24], cov=0.5, G=3, E=4, T=40, restart=False, seed=25, gamma=0.5,
                            eps_clip=0.15),
        "pendulum_h0": dict(kind=1, hidden=[], cov=0.5, G=3, E=4, T=30, restart=True, seed=26, gamma=0.99, eps_clip=0.2),
        "cartpole_tanh": dict(kind=0, hidden=[32, 32], cov=0.5, G=3, E=4, T=40, restart=False, seed=27, gamma=0.5,
                              eps_clip=0.15, activation="Tanh"),
        "pendulum_mixed_act": dict(kind=1, hidden=[32, 24], cov=0.5, G=3, E=4, T=30, restart=True, seed=28, gamma=0.99,
                                   eps_clip=0.2, activation=["Tanh", "ReLU"]),
    }
    for name, kw in roll.items():
        F[f"rollout_grpo_{name}"] = (lambda kw=kw: gen_rollout_and_learn(**kw))

    ppo = dict(kind=2, hidden=[32, 32], cov=0.5, G=2, E=3, T=40, seed=7, gamma=0.99, lam=0.95, eps_clip=0.2)
    F["ppo_mc_quadpole2d"] = lambda: gen_ppo(monte_carlo=True, **ppo)
    F["ppo_gae_quadpole2d"] = lambda: gen_ppo(monte_carlo=False, **ppo)
    # ragged full batch (T = 120: episodes end by leaving the box at different steps)
    rag = dict(kind=2, hidden=[32, 32], cov=0.5, G=3, E=4, T=120, seed=12, gamma=0.99, lam=0.95, eps_clip=0.2)
    F["ppo_mc_ragged_quadpole2d"] = lambda: gen_ppo(monte_carlo=True, **rag)
    F["ppo_gae_ragged_quadpole2d"] = lambda: gen_ppo(monte_carlo=False, **rag)
    # the shipped QuadPole2D PPO checkpoint: actor + critic 10-128-128-128-{2,1} (three hidden layers), trained
    # weights => long episodes (pipelines/quadpole2d_pipeline_ppo.py:54-80; GAE off = Monte-Carlo is its default)
    F["ppo_mc_quadpole2d_shipped"] = lambda: gen_ppo(kind=2, hidden=[128, 128, 128], cov=0.5, G=2, E=3, T=150, seed=13,
                                                     gamma=0.99, lam=0.95, eps_clip=0.2, monte_carlo=True, weights=QP2_PPO)
    F["ppo_gae_cartpole_shipped"] = lambda: gen_ppo(kind=0, hidden=[128, 128, 128], cov=0.5, G=2, E=3, T=120, seed=14,
                                                    gamma=0.99, lam=0.95, eps_clip=0.2, monte_carlo=False, weights=CP_PPO)
    F["ppo_minibatch_quadpole2d"] = lambda: gen_ppo_minibatch(kind=2, hidden=[32, 32], cov=0.5, G=3, E=4, T=120, seed=11,
                                                              gamma=0.99, lam=0.95, eps_clip=0.2, batch_size=64, updates=2,
                                                              torch_seed=4321)
    # resume from the shipped checkpoints (Pipeline(load_path=...) semantics), then one learn()
    F["resume_grpo_cartpole"] = lambda: gen_resume("grpo", 0, [128, 128, 128, 128], SHIPPED[0], G=3, E=4, T=60, seed=31,
                                                   updates=2)
    F["resume_ppo_cartpole"] = lambda: gen_resume("ppo", 0, [128, 128, 128], SHIPPED[1], G=2, E=3, T=80, seed=32, updates=2)
    F["resume_ppo_quadpole2d"] = lambda: gen_resume("ppo", 2, [128, 128, 128], SHIPPED[2], G=2, E=3, T=80, seed=33,
                                                    updates=2)
    return F


def main():
    """python oracle/make_golden.py [--only a,b] [--check]"""
    if not ref_shims.available():
        raise SystemExit("reference not mounted; fixtures can only be regenerated in the build container")
    ref_shims.install()
    os.makedirs(OUT, exist_ok=True)
    check = "--check" in sys.argv
    if not check:
        copy_shipped()
    only = None
    if "--only" in sys.argv:
        only = set(sys.argv[sys.argv.index("--only") + 1].split(","))
    report = {}
    for i, (name, gen) in enumerate(fixtures().items()):
        if only is not None and name not in only:
            continue
        np.random.seed(777000 + i)       # numpy's global RNG: Env.reset() (e.g. cartpole_env.py:103) draws from it
        torch.manual_seed(888000 + i)
        arrays = gen()
        emit(name, arrays, check, report)
        if "len" in arrays:
            ln = np.asarray(arrays["len"]).reshape(-1)
            print(f"  {name}: lens min/mean/max = {ln.min():.0f}/{ln.mean():.1f}/{ln.max():.0f}")
        elif "done" in arrays:
            print(f"  {name}: {len(arrays['reward'])} rows, done={int(arrays['done'].sum())}")
    if check and any(report.values()):
        raise SystemExit("fixtures do not regenerate bit-identically: " + str({k: v for k, v in report.items() if v}))


if __name__ == "__main__":
    main()
