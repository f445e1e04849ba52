"""The reference-facing host classes (same names and protocol as the reference)
driven end to end on the GPU and checked against the fixtures generated from the
unmodified reference."""
import os

import numpy as np
import pytest
import torch

import restate as R

pytestmark = pytest.mark.gpu


def load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name), allow_pickle=False))


@pytest.fixture(scope="module")
def tg():
    import trajopt_grpo_b200 as pkg
    return pkg


ENV_CLS = {0: "CartPole", 1: "Pendulum", 2: "QuadPole2D", 3: "QuadPole"}
MAXS = {0: 120, 1: 200, 2: 150, 3: 150}
SPLIT = {0: {"cartpole": 5}, 1: {"pendulum": 3}, 2: {"quadrotor": 7, "pendulum": 3}, 3: {"quadrotor": 13, "pendulum": 7}}


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_env_step_protocol_matches_reference(tg, golden_dir, kind):
    """Env.reset/restart/step with state injection (the pattern of the reference's
    tests/test_cartpole.py:94-98) reproduces recorded reference transitions."""
    g = load(golden_dir, f"transitions_env{kind}.npz")
    env = getattr(tg, ENV_CLS[kind])(max_steps=MAXS[kind])
    obs, info = env.reset()
    assert obs.shape == env.observation_space.shape and "time_balanced" in info
    # first episode of the fixture: consecutive transitions from steps_done == 0
    idx = [0]
    while idx[-1] + 1 < len(g["done"]) and not g["done"][idx[-1]] and len(idx) < 25:
        idx.append(idx[-1] + 1)
    off = 0
    init = {}
    for k, n in SPLIT[kind].items():
        init[k] = g["state"][0][off:off + n].copy(); off += n
    env._initial_state = init
    obs, _ = env.restart()
    np.testing.assert_allclose(obs, g["state"][0], rtol=0, atol=0)
    for i in idx:
        obs, rew, f1, f2, info = env.step(g["action"][i])
        np.testing.assert_allclose(obs, g["next"][i], rtol=1e-9, atol=1e-10)
        assert abs(rew - g["reward"][i]) <= 1e-9 * max(1.0, abs(g["reward"][i]))
        assert bool(f1 or f2) == bool(g["done"][i])


def _weights(g, prefix=""):
    Ws, bs, i = [], [], 0
    while f"{prefix}W{i}" in g:
        Ws.append(g[f"{prefix}W{i}"]); bs.append(g[f"{prefix}b{i}"]); i += 1
    return Ws, bs


def _load_actor(net, Ws, bs):
    sd = {}
    for i, (w, b) in enumerate(zip(Ws, bs)):
        sd[f"network.{2 * i}.weight"] = torch.from_numpy(w)
        sd[f"network.{2 * i}.bias"] = torch.from_numpy(b)
    net.load_state_dict(sd)          # the reference's state-dict keys (models/neural_network.py:50-65)


class _Buf:
    pass


def _make_buffer(tg, g):
    buf = tg.Rollout_Buffer.__new__(tg.Rollout_Buffer)
    buf.avg_reward = []
    buf.device_rollout = None
    tg.Rollout_Buffer.store(buf, g["obs"], g["act"], g["rew"], g["len"], g["mask"])
    return buf


ROLLOUTS = ["cartpole", "pendulum", "quadpole2d", "quadpole"]


@pytest.mark.parametrize("name", ROLLOUTS)
def test_grpo_learn_sgd_and_adam_match_reference(tg, golden_dir, name):
    g = load(golden_dir, f"rollout_grpo_{name}.npz")
    kind = int(g["kind"])
    Ws, bs = _weights(g)
    hidden = [int(h) for h in g["hidden"]]
    O, A = R.OBS_DIM[kind], R.ACT_DIM[kind]
    buf = _make_buffer(tg, g)
    assert abs(float(buf.avg_reward[-1]) - float(g["rew"].sum(2).mean())) < 1e-3
    # --- SGD(lr=1), one update: theta_before - theta_after = gradient (generic optimizer path)
    pol = tg.GaussianActor_NeuralNetwork(O, A, hidden, "ReLU", float(g["cov"]))
    _load_actor(pol.actor, Ws, bs)
    before = [p.detach().clone() for p in pol.parameters()]
    algo = tg.GRPO(float(g["eps_clip"]), 0.0, float(g["gamma"]), pol, torch.optim.SGD(pol.parameters(), lr=1.0), None,
                   updates_per_iter=1)
    algo.learn(buf)
    for i, (b, p) in enumerate(zip(before, pol.parameters())):
        ref = g[f"grpo_grad{i}"]
        got = (b - p.detach()).cpu().numpy()
        assert np.abs(got - ref).max() <= 3e-4 * max(np.abs(ref).max(), 1e-6) + 1e-5
    # --- Adam, 3 updates, twice (fused tg_adam_step path, old-policy sync in between)
    pol = tg.GaussianActor_NeuralNetwork(O, A, hidden, "ReLU", float(g["cov"]))
    _load_actor(pol.actor, Ws, bs)
    opt = torch.optim.Adam(pol.parameters(), lr=3e-4)
    algo = tg.GRPO(float(g["eps_clip"]), 0.0, float(g["gamma"]), pol, opt, None, updates_per_iter=3)
    for key in ("grpo_adam3_p", "grpo_adam6_p"):
        algo.learn(buf)
        for i, p in enumerate(pol.parameters()):
            np.testing.assert_allclose(p.detach().cpu().numpy(), g[f"{key}{i}"], rtol=2e-4, atol=3e-6)
    # optimizer state is checkpointable in torch's own format (grpo.py:150-160)
    sd = opt.state_dict()
    assert int(float(sd["state"][0]["step"])) == 6 and sd["state"][0]["exp_avg"].shape == Ws[0].shape
    for a, b in zip(algo.old_policy.parameters(), pol.parameters()):
        assert torch.equal(a, b)                                     # grpo.py:148


@pytest.mark.parametrize("name", ["mc", "gae"])
def test_ppo_learn_matches_reference(tg, golden_dir, name):
    g = load(golden_dir, f"ppo_{name}_quadpole2d.npz")
    kind = int(g["kind"])
    hidden = [int(h) for h in g["hidden"]]
    O, A = R.OBS_DIM[kind], R.ACT_DIM[kind]
    Ws, bs = _weights(g)
    cWs, cbs = _weights(g, "c")
    buf = _make_buffer(tg, g)
    kw = dict(c1=0.5, kl_coeff=0.5, gamma=float(g["gamma"]), lam=float(g["lam"]), entropy=0.01, batch_size=None,
              monte_carlo=bool(g["monte_carlo"]))

    def fresh():
        pol = tg.GaussianActorCritic_NeuralNetwork(O, A, hidden, "ReLU", float(g["cov"]))
        _load_actor(pol.actor, Ws, bs)
        _load_actor(pol.critic, cWs, cbs)
        return pol

    pol = fresh()
    before = [p.detach().clone() for p in pol.parameters()]
    tg.PPO(float(g["eps_clip"]), pol, torch.optim.SGD(pol.parameters(), lr=1.0), None, 1, **kw).learn(buf)
    for i, (b, p) in enumerate(zip(before, pol.parameters())):
        ref = g[f"ppo_grad{i}"]
        got = (b - p.detach()).cpu().numpy()
        assert np.abs(got - ref).max() <= 5e-4 * max(np.abs(ref).max(), 1e-6) + 1e-6, (i, np.abs(got - ref).max())
    pol = fresh()
    tg.PPO(float(g["eps_clip"]), pol, torch.optim.Adam(pol.parameters(), lr=2e-4), None, 3, **kw).learn(buf)
    for i, p in enumerate(pol.parameters()):
        np.testing.assert_allclose(p.detach().cpu().numpy(), g[f"ppo_adam3_p{i}"], rtol=5e-4, atol=5e-6)


def test_rollout_manager_and_buffer_protocol(tg, tmp_path):
    torch.manual_seed(0)
    pol = tg.GaussianActor_NeuralNetwork(3, 1, [64, 64], "ReLU", 0.5)
    mgr = tg.RolloutManager(lambda: tg.Pendulum(max_steps=50), pol, restart=True, num_workers=6,
                            num_episodes_per_worker=16, use_multiprocessing=False, seed=3)
    obs, act, rew, ln, mask = mgr.rollout()
    assert obs.shape == (6, 16, 50, 3) and act.shape == (6, 16, 50, 1) and rew.shape == (6, 16, 50)
    assert ln.shape == (6, 16) and ln.dtype == torch.float32 and mask.shape == (6, 16, 50)
    assert torch.equal(mask.sum(2), ln)
    # restart=True: the E episodes of a group share their initial observation (rollout_worker.py:70-71)
    assert torch.equal(obs[:, :1, 0].expand(-1, 16, -1), obs[:, :, 0])
    assert not torch.equal(obs[0, :, 0], obs[1, :, 0])
    # same seed -> same rollout; the stream advances between rollouts
    mgr2 = tg.RolloutManager(lambda: tg.Pendulum(max_steps=50), pol, restart=True, num_workers=6,
                             num_episodes_per_worker=16, seed=3)
    assert torch.equal(mgr2.rollout()[0], obs)
    assert not torch.equal(mgr2.rollout()[1], act)
    buf = tg.Rollout_Buffer(mgr)
    buf.sample()
    assert buf.group_observations.shape == (6, 16, 50, 3) and len(buf.avg_reward) == 1
    assert abs(float(buf.avg_reward[0]) - float(buf.group_rewards.sum(2).mean())) < 1e-3
    # Dashboard-style indexing (visualize/visualizer.py:120-134)
    i, ep = 1, 2
    n = int(buf.group_lengths[i, ep])
    assert buf.group_observations[i, ep, n - 1].shape == (3,)
    # the render feed: only the first k episodes of the first 4 groups travel to the host
    hs = buf.host_slice(4, 5)
    assert hs["observations"].shape == (4, 5, 50, 3) and hs["lengths"].shape == (4, 5)
    np.testing.assert_array_equal(hs["observations"], buf.group_observations[:4, :5].cpu().numpy())
    np.testing.assert_array_equal(hs["actions"], buf.group_actions[:4, :5].cpu().numpy())
    np.testing.assert_array_equal(hs["lengths"], buf.group_lengths[:4, :5].cpu().numpy().astype(int))
    # one training epoch through the reference's two calls (pipelines/pipeline.py:163-164)
    opt = torch.optim.Adam(pol.parameters(), lr=5e-4)
    algo = tg.GRPO(0.2, 0.01, 0.99, pol, opt, None, updates_per_iter=2)
    w0 = pol.flat_parameters().clone()
    algo.learn(buf)
    assert not torch.equal(w0, pol.flat_parameters()) and bool(torch.isfinite(pol.flat_parameters()).all())
    # checkpoint round trip in the reference's file formats (pipeline.py:104-118)
    d = str(tmp_path)
    pol.save(d); algo.save(d); buf.save(d); buf.save_trajectory(d)
    sd = torch.load(os.path.join(d, "policy.pt"), weights_only=True)
    assert list(sd) == ["network.0.weight", "network.0.bias", "network.2.weight", "network.2.bias",
                        "network.4.weight", "network.4.bias"]
    pol2 = tg.GaussianActor_NeuralNetwork(3, 1, [64, 64], "ReLU", 0.5)
    pol2.load(d)
    assert torch.equal(pol2.flat_parameters(), pol.flat_parameters())
    algo2 = tg.GRPO(0.2, 0.01, 0.99, pol2, torch.optim.Adam(pol2.parameters(), lr=5e-4), None, updates_per_iter=2)
    algo2.load(d)
    algo.learn(buf); algo2.learn(buf)
    assert torch.allclose(pol2.flat_parameters(), pol.flat_parameters(), rtol=0, atol=0)
    assert buf.load(d) == 1
    import pandas as pd
    df = pd.read_csv(os.path.join(d, "trajectory.csv"))
    assert len(df) == int(buf.group_lengths.sum()) and list(df.columns[:2]) == ["episode_id", "observation_0"]


def test_policy_forward_and_log_prob_api(tg):
    torch.manual_seed(1)
    pol = tg.GaussianActorCritic_NeuralNetwork(10, 2, [128, 128], "ReLU", 0.5)
    st = np.random.default_rng(0).standard_normal(10)
    a, lp, v = pol(st)
    assert isinstance(a, np.ndarray) and a.shape == (2,) and a.dtype == np.float32 and v.shape == (1,)
    obs = torch.randn(33, 10)
    act = torch.randn(33, 2)
    lp, ent = pol.log_prob(obs, act)
    Ws = [pol.actor.network[i].weight.detach().cpu().numpy() for i in (0, 2, 4)]
    bs = [pol.actor.network[i].bias.detach().cpu().numpy() for i in (0, 2, 4)]
    mu = R.mlp_forward(obs.numpy(), Ws, bs, 0, np.float64)
    ref = R.gaussian_log_prob(mu, np.array([0.5, 0.5], np.float32), act.numpy(), np.float64)
    np.testing.assert_allclose(lp.cpu().numpy(), ref, rtol=1e-5, atol=2e-5)
    assert abs(float(ent[0]) - R.gaussian_entropy(np.array([0.5, 0.5]), 2)) < 1e-6
    assert pol.value(obs).shape == (33,)
    assert pol.metadata()["num_parameters"] == sum(p.numel() for p in pol.parameters())


def test_quadrotor_dynamics_api(tg, golden_dir):
    g = load(golden_dir, "quadrotor12_dynamics.npz")
    q = tg.Quadrotor()
    out = q._dynamics(g["state"][0], g["control"][0])
    np.testing.assert_allclose(out, g["next"][0], rtol=1e-11, atol=1e-12)


def test_trajectory_export_matches_reference_loop(tg, tmp_path):
    """Rollout_Buffer.save_trajectory (rollout_buffer.py:72-102) through the device compaction kernel
    against the reference's host loop restated on the [G,E,T,.] views; ragged episodes."""
    import pandas as pd
    torch.manual_seed(2)
    pol = tg.GaussianActor_NeuralNetwork(10, 2, [32, 32], "ReLU", 0.5)
    mgr = tg.RolloutManager(lambda: tg.QuadPole2D(max_steps=60), pol, restart=False, num_workers=5,
                            num_episodes_per_worker=7, use_multiprocessing=False, seed=9)
    buf = tg.Rollout_Buffer(mgr)
    buf.sample()
    ln = buf.group_lengths.cpu().numpy().astype(int)
    assert ln.min() < ln.max()                      # the fixture must be ragged
    obs, act = buf.group_observations.cpu().numpy(), buf.group_actions.cpu().numpy()
    rows, ids = [], []
    for i in range(ln.shape[0]):
        for j in range(ln.shape[1]):
            rows.append(np.hstack([obs[i, j, :ln[i, j]], act[i, j, :ln[i, j]]]))
            ids.extend([j + i * ln.shape[1]] * ln[i, j])
    want = np.vstack(rows).astype(np.float64)
    buf.save_trajectory(str(tmp_path))
    df = pd.read_csv(os.path.join(str(tmp_path), "trajectory.csv"))
    assert list(df.columns) == ["episode_id"] + [f"observation_{i}" for i in range(10)] + ["action_0", "action_1"]
    assert df["episode_id"].tolist() == ids
    # the CSV text round trip of a float64 is good to ~1e-15; the table itself is bit-identical
    np.testing.assert_allclose(df.values[:, 1:], want, rtol=1e-12, atol=0)


def _epoch_loop(policy, algo, buf, epochs, save_dir):
    """The calls the reference's orchestration makes on the hot-path objects, and nothing else
    (pipelines/pipeline.py:163-164 sample -> learn; :111-113 the three save() calls).  The reference's own
    Pipeline class drives the engine's classes unchanged; tests/test_dropin_surface.py checks in the build
    container that every attribute it touches exists here."""
    for _ in range(epochs):
        buf.sample()
        algo.learn(buf)
        for part in (algo, policy, buf):
            part.save(save_dir)


def test_train_checkpoint_resume_through_the_reference_protocol(tg, tmp_path):
    """train -> checkpoint in the reference's file formats -> fresh objects load() it (what
    Pipeline(load_path=...) does, pipeline.py:93-102) -> identical policy, optimizer state and reward history,
    and the resumed run continues bit-identically to the uninterrupted one."""
    d = str(tmp_path)

    def build():
        torch.manual_seed(4)
        pol = tg.GaussianActor_NeuralNetwork(5, 1, [32, 32], "ReLU", 0.5)
        opt = torch.optim.Adam(pol.parameters(), lr=3e-4)
        algo = tg.GRPO(0.15, 0.5, 0.5, pol, opt, None, updates_per_iter=1)
        mgr = tg.RolloutManager(lambda: tg.CartPole(max_steps=40), pol, restart=False, num_workers=4,
                                num_episodes_per_worker=5, use_multiprocessing=False, seed=1)
        return pol, opt, algo, mgr, tg.Rollout_Buffer(mgr)

    pol, opt, algo, mgr, buf = build()
    _epoch_loop(pol, algo, buf, 3, d)
    for f in ("policy.pt", "optimizer.pth", "reward.csv"):
        assert os.path.exists(os.path.join(d, f)), f
    meta = {"policy": pol.metadata(), "algorithm": algo.metadata(), "buffer": buf.metadata()}     # pipeline.py:120-139
    assert meta["policy"]["hidden_dims"] == [32, 32]
    assert meta["algorithm"] == {"algorithm": "GRPO", "epsilon": 0.15, "beta": 0.5, "updates_per_iter": 1}
    assert meta["policy"]["num_parameters"] == 5 * 32 + 32 + 32 * 32 + 32 + 32 + 1
    assert isinstance(meta["buffer"]["avg_reward"], float)
    pol2, opt2, algo2, mgr2, buf2 = build()
    algo2.load(d); pol2.load(d); n = buf2.load(d)                  # pipeline.py:98-100, in its order
    assert n == 3 and np.allclose(buf2.avg_reward, [float(x) for x in buf.avg_reward])
    assert torch.equal(pol2.flat_parameters(), pol.flat_parameters())
    # continue both: same manager stream position is part of the resume contract of OUR seedable manager
    mgr2._epoch, mgr2._rng = mgr._epoch, np.random.default_rng(123)
    mgr._rng = np.random.default_rng(123)
    buf.sample(); algo.learn(buf)
    buf2.sample(); algo2.learn(buf2)
    assert torch.equal(pol2.flat_parameters(), pol.flat_parameters())
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert s1.keys() == s2.keys()
    for k in s1:
        assert torch.equal(s1[k]["exp_avg"].cpu(), s2[k]["exp_avg"].cpu())
        assert float(s1[k]["step"]) == float(s2[k]["step"]) == 4.0
    buf2.save_trajectory(d)
    assert os.path.exists(os.path.join(d, "trajectory.csv"))
    mgr.shutdown(); mgr2.shutdown()                                # pipeline.py:207


def test_learn_sees_weights_loaded_after_construction(tg, golden_dir):
    """weights written through torch (load_state_dict, p.add_) after GRPO was constructed, or into
    old_policy, must defeat the 'rollout log-prob == old-policy log-prob' shortcut (grpo.py:118-119)."""
    g = load(golden_dir, "rollout_grpo_pendulum.npz")
    Ws, bs = _weights(g)
    hidden = [int(h) for h in g["hidden"]]
    torch.manual_seed(0)
    pol = tg.GaussianActor_NeuralNetwork(3, 1, hidden, "ReLU", float(g["cov"]))
    algo = tg.GRPO(float(g["eps_clip"]), 0.0, float(g["gamma"]), pol, torch.optim.Adam(pol.parameters(), lr=3e-4), None,
                   updates_per_iter=3)
    t0 = pol.param_tag()
    _load_actor(pol.actor, Ws, bs)                                 # load AFTER construction: old_policy is stale
    assert pol.param_tag() != t0
    with torch.no_grad():
        next(iter(pol.parameters())).add_(0.0)
    mgr = tg.RolloutManager(lambda: tg.Pendulum(max_steps=int(g["T"])), pol, restart=True, num_workers=int(g["G"]),
                            num_episodes_per_worker=int(g["E"]), use_multiprocessing=False, seed=0)
    buf = tg.Rollout_Buffer(mgr)
    init = torch.from_numpy(g["init"].T.copy()).float().cuda()
    nz = torch.from_numpy(np.ascontiguousarray(g["noise"].transpose(0, 2, 1))).cuda()
    buf.sample(init_state=init, noise=nz)
    # reference semantics: old log-probs come from old_policy (random init here), NOT from the rollout weights
    old = algo.old_policy
    ref_pol = tg.GaussianActor_NeuralNetwork(3, 1, hidden, "ReLU", float(g["cov"]))
    ref_pol.load_state_dict(pol.state_dict())
    ref_algo = tg.GRPO(float(g["eps_clip"]), 0.0, float(g["gamma"]), ref_pol, torch.optim.Adam(ref_pol.parameters(), lr=3e-4),
                       None, updates_per_iter=3)
    ref_algo.old_policy.load_state_dict(old.state_dict())          # explicit stale old policy: never takes the shortcut
    assert ref_algo.old_policy.param_tag() != ref_algo._old_tag
    algo.learn(buf)
    ref_algo.learn(buf)
    assert torch.equal(pol.flat_parameters(), ref_pol.flat_parameters())
    # and the shortcut's own result differs (ratio != 1 from the first update), so the check above is not vacuous
    p3 = tg.GaussianActor_NeuralNetwork(3, 1, hidden, "ReLU", float(g["cov"]))
    _load_actor(p3.actor, Ws, bs)                                  # loaded BEFORE GRPO copies it: old == current
    a3 = tg.GRPO(float(g["eps_clip"]), 0.0, float(g["gamma"]), p3, torch.optim.Adam(p3.parameters(), lr=3e-4), None,
                 updates_per_iter=3)
    mgr3 = tg.RolloutManager(lambda: tg.Pendulum(max_steps=int(g["T"])), p3, restart=True, num_workers=int(g["G"]),
                             num_episodes_per_worker=int(g["E"]), use_multiprocessing=False, seed=0)
    b3 = tg.Rollout_Buffer(mgr3)
    b3.sample(init_state=init, noise=nz)
    a3.learn(b3)
    assert not torch.equal(p3.flat_parameters(), pol.flat_parameters())


RESUME = [("resume_grpo_cartpole", "grpo"), ("resume_ppo_cartpole", "ppo"), ("resume_ppo_quadpole2d", "ppo")]


@pytest.mark.parametrize("name,kind_name", RESUME)
def test_resume_from_shipped_checkpoints(tg, golden_dir, name, kind_name):
    """The three checkpoints the reference ships under reports/** (copied as data to tests/golden/shipped by
    oracle/make_golden.py) load through algorithm.load / policy.load / buffer.load, and one learn() from the
    loaded Adam state reproduces the unmodified reference's resumed update."""
    import json
    g = load(golden_dir, f"{name}.npz")
    path = os.path.join(golden_dir, "shipped", str(g["report"]))
    meta = json.load(open(os.path.join(path, "metadata.json")))
    kind = int(g["kind"])
    hidden = [int(h) for h in g["hidden"]]
    assert hidden == meta["policy"]["hidden_dims"]
    O, A = R.OBS_DIM[kind], R.ACT_DIM[kind]
    am = meta["algorithm"]
    torch.manual_seed(0)
    if kind_name == "ppo":
        pol = tg.GaussianActorCritic_NeuralNetwork(O, A, hidden, meta["policy"]["activation"], 0.5)
        opt = torch.optim.Adam(pol.parameters(), lr=2e-4)
        algo = tg.PPO(am["epsilon"], pol, opt, None, int(g["updates"]), c1=am["c1"], kl_coeff=am["kl_coeff"],
                      gamma=am["gamma"], lam=am["lam"], entropy=am["entropy"], batch_size=am["batch_size"], monte_carlo=True)
    else:
        pol = tg.GaussianActor_NeuralNetwork(O, A, hidden, meta["policy"]["activation"], 0.5)
        opt = torch.optim.Adam(pol.parameters(), lr=3e-4)
        algo = tg.GRPO(am["epsilon"], am["beta"], 0.5, pol, opt, None, updates_per_iter=int(g["updates"]))
    mgr = tg.RolloutManager(lambda: getattr(tg, ENV_CLS[kind])(max_steps=int(g["T"])), pol, num_workers=int(g["G"]),
                            num_episodes_per_worker=int(g["E"]), use_multiprocessing=False, seed=0)
    buf = tg.Rollout_Buffer(mgr)
    algo.load(path); pol.load(path); n_epochs = buf.load(path)      # pipelines/pipeline.py:98-100
    assert n_epochs == int(g["n_epochs_loaded"])
    assert pol.metadata()["num_parameters"] == meta["policy"]["num_parameters"]
    sd = torch.load(os.path.join(path, "policy.pt"), weights_only=True)
    first = (sd["actor"] if kind_name == "ppo" else sd)["network.0.weight"]
    assert torch.equal(next(iter(pol.parameters())).detach().cpu(), first)
    tg.Rollout_Buffer.store(buf, g["obs"], g["act"], g["rew"], g["len"], g["mask"])
    algo.learn(buf)
    assert algo._flat_opt.step_count == int(g["step_after"]) == int(g["step_loaded"]) + int(g["updates"])
    for i, p in enumerate(pol.parameters()):
        np.testing.assert_allclose(p.detach().cpu().numpy(), g[f"resume_p{i}"], rtol=5e-4, atol=5e-6, err_msg=f"param {i}")
    assert int(float(opt.state_dict()["state"][0]["step"])) == int(g["step_after"])


@pytest.mark.parametrize("name,kind,keys", [("cartpole", 0, ("masscart", "masspole", "length", "gravity")),
                                            ("pendulum", 1, ("mass", "length", "gravity"))])
def test_env_with_non_default_physical_parameters(tg, golden_dir, name, kind, keys):
    """CartPole / Pendulum constructed with non-default masses, length, gravity, timestep (cartpole_env.py:7-16,
    pendulum_env.py:8-17): batched tg_env_step in float64 vs the reference's transitions, Env._dynamics, and a
    free-running rollout against the oracle with the same parameters."""
    from trajopt_grpo_b200 import engine as E
    g = load(golden_dir, f"transitions_{name}_params.npz")
    kw = {k: float(g[k]) for k in keys}
    dt = float(g["timestep"])
    env = getattr(tg, ENV_CLS[kind])(max_steps=100, timestep=dt, **kw)
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    nxt, rew, done, _ = E.env_step(kind, 100, dt, dev(g["state"].T), dev(g["action"].T), dev(g["steps_done"].astype(np.int32)),
                                   dev(g["bal_count"].astype(np.int32)), phys=env._tg_phys)
    np.testing.assert_allclose(nxt.cpu().numpy().T, g["next"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(rew.cpu().numpy(), g["reward"], rtol=1e-10, atol=1e-11)
    assert np.array_equal(done.cpu().numpy().astype(bool), g["done"])
    for i in (0, 7, 50):
        np.testing.assert_allclose(env._dynamics(g["state"][i], g["control"][i]), g["dyn_next"][i], rtol=1e-11, atol=1e-12)
    # free-running rollout through the host classes vs the oracle with the same physical parameters
    rng = np.random.default_rng(3)
    torch.manual_seed(0)
    pol = tg.GaussianActor_NeuralNetwork(R.OBS_DIM[kind], 1, [32, 32], "ReLU", 0.5)
    Gn, En, T = 4, 4, 25
    mgr = tg.RolloutManager(lambda: getattr(tg, ENV_CLS[kind])(max_steps=T, timestep=dt, **kw), pol, num_workers=Gn,
                            num_episodes_per_worker=En, use_multiprocessing=False, seed=0, precision="f64")
    init = R.reset_states(kind, Gn * En, rng)
    noise = rng.standard_normal((T, Gn * En, 1)).astype(np.float32)
    r = mgr.rollout_device(init_state=dev(init.T), noise=dev(noise.transpose(0, 2, 1)))
    Ws = [pol.actor.network[i].weight.detach().cpu().numpy() for i in (0, 2, 4)]
    bs = [pol.actor.network[i].bias.detach().cpu().numpy() for i in (0, 2, 4)]
    cfg = R.EnvCfg.make(kind, T, dt, phys=[kw[k] for k in keys])
    o, a, rw, lp, ln, m = R.rollout(cfg, init, Ws, bs, np.array([0.5], np.float32), noise)
    assert np.array_equal(r.len.cpu().numpy(), ln)
    np.testing.assert_allclose(r.obs.cpu().numpy().transpose(2, 0, 1), o, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(r.rew.cpu().numpy().T, rw, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("name,kind", [("dynamics_quadpole2d", 2), ("dynamics_quadpole", 3), ("transitions_cartpole_params", None)])
def test_env_dynamics_matches_reference(tg, golden_dir, name, kind):
    """Env._dynamics(state, control) (quadrotor_env.py:417-528, 1044-1130) through tg_env_dynamics: float64 vs the
    reference's direct _dynamics calls, float32 within the throughput-mode tolerance."""
    from trajopt_grpo_b200 import engine as E
    if kind is None:
        return
    g = load(golden_dir, f"{name}.npz")
    dev = lambda x, dt=None: torch.from_numpy(np.ascontiguousarray(x)).cuda() if dt is None else \
        torch.from_numpy(np.ascontiguousarray(x)).cuda().to(dt)
    nxt = E.env_dynamics(kind, 0.02, dev(g["state"].T), dev(g["control"].T))
    np.testing.assert_allclose(nxt.cpu().numpy().T, g["dyn_next"], rtol=1e-11, atol=1e-12)
    nxt32 = E.env_dynamics(kind, 0.02, dev(g["state"].T, torch.float32), dev(g["control"].T))
    np.testing.assert_allclose(nxt32.cpu().numpy().T, g["dyn_next"], rtol=1e-5, atol=1e-5)
    env = getattr(tg, ENV_CLS[kind])()
    np.testing.assert_allclose(env._dynamics(g["state"][3], g["control"][3]), g["dyn_next"][3], rtol=1e-11, atol=1e-12)


def test_ppo_minibatched_learn_matches_reference(tg, golden_dir):
    """PPO.learn with batch_size=64 (the reference constructor's default, ppo.py:147-183): randperm
    minibatches over the valid steps of a ragged rollout, one Adam step per minibatch.  The fixture was
    produced by the unmodified reference with torch's CPU generator seeded right before learn()."""
    g = load(golden_dir, "ppo_minibatch_quadpole2d.npz")
    kind = int(g["kind"])
    hidden = [int(h) for h in g["hidden"]]
    O, A = R.OBS_DIM[kind], R.ACT_DIM[kind]
    Ws, bs = _weights(g)
    cWs, cbs = _weights(g, "c")
    buf = _make_buffer(tg, g)
    assert len(set(g["len"].reshape(-1).tolist())) > 1          # ragged
    pol = tg.GaussianActorCritic_NeuralNetwork(O, A, hidden, "ReLU", float(g["cov"]))
    _load_actor(pol.actor, Ws, bs)
    _load_actor(pol.critic, cWs, cbs)
    algo = tg.PPO(float(g["eps_clip"]), pol, torch.optim.Adam(pol.parameters(), lr=2e-4), None, int(g["updates"]),
                  c1=0.5, kl_coeff=0.5, gamma=float(g["gamma"]), lam=float(g["lam"]), entropy=0.01,
                  batch_size=int(g["batch_size"]), monte_carlo=False)
    torch.manual_seed(int(g["torch_seed"]))
    algo.learn(buf)
    n_steps = int(g["updates"]) * -(-int(g["len"].sum()) // int(g["batch_size"]))
    assert algo._flat_opt.step_count == n_steps                   # one optimizer step per minibatch
    for i, p in enumerate(pol.parameters()):
        np.testing.assert_allclose(p.detach().cpu().numpy(), g[f"ppo_mb_adam_p{i}"], rtol=1e-3, atol=2e-5)


@pytest.mark.parametrize("updates", [1, 3])
def test_streamed_epoch_equals_sample_plus_learn(tg, updates):
    """GRPO.learn_streamed (BASELINE configs[4]: more envs per GPU than trajectories fit in HBM) keeps only the
    initial states and the Philox seed and rematerialises the rollout chunk by chunk (whole groups) in every
    update; it must train the same weights as sample() + learn() on the same seed.  The only difference is the
    fp32 order in which the per-chunk gradients are summed."""
    def build():
        torch.manual_seed(5)
        pol = tg.GaussianActor_NeuralNetwork(20, 4, [256, 256], "ReLU", 0.3)
        opt = torch.optim.Adam(pol.parameters(), lr=3e-4)
        algo = tg.GRPO(0.2, 0.01, 0.999, pol, opt, None, updates_per_iter=updates)
        mgr = tg.RolloutManager(lambda: tg.QuadPole(max_steps=60), pol, restart=True, num_workers=12,
                                num_episodes_per_worker=16, use_multiprocessing=False, seed=11)
        return pol, algo, mgr

    rng = np.random.default_rng(2)
    env = tg.QuadPole(max_steps=60)
    s0 = np.repeat(env.sample_initial_states(12, rng), 16, axis=0)
    init = torch.from_numpy(np.ascontiguousarray(s0.T)).float().cuda()
    pol_a, algo_a, mgr_a = build()
    buf = tg.Rollout_Buffer(mgr_a)
    buf.sample(init_state=init)
    n_ref = int(buf.device_rollout.len.sum())
    algo_a.learn(buf)
    pol_b, algo_b, mgr_b = build()
    n_valid = algo_b.learn_streamed(mgr_b, init_state=init, chunk_groups=5)      # chunks of 5, 5 and 2 groups
    assert int(n_valid) == n_ref
    assert abs(float(algo_b.last_mean_return) - float(buf.avg_reward[-1])) < 1e-3 * max(1.0, abs(float(buf.avg_reward[-1])))
    a, b = pol_a.flat_parameters(), pol_b.flat_parameters()
    # Adam's first steps move every element by ~lr; a summation-order difference of 1e-7 relative in the gradient
    # cannot flip a step except for elements whose gradient is ~0: bound by 2 lr per update, 99.9 % within 1e-6
    d = (a - b).abs()
    assert float(d.max()) <= 2 * 3e-4 * updates + 1e-7
    assert float((d <= 1e-6).float().mean()) >= 0.999
    # second epoch through the streamed path continues from synchronised old-policy weights
    algo_b.learn_streamed(mgr_b, init_state=init, chunk_groups=12)
    assert bool(torch.isfinite(pol_b.flat_parameters()).all())
