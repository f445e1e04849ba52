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


def test_pipeline_train_save_load_resume(tg, tmp_path):
    """pipelines/pipeline.py: train -> archive checkpoint in the reference's file formats -> a new
    Pipeline(load_path=...) resumes with identical policy, optimizer state and reward history."""
    import json

    def build(load_path=None):
        torch.manual_seed(4)
        pol = tg.GaussianActor_NeuralNetwork(5, 1, [32, 32], "ReLU", 0.5)
        opt = torch.optim.Adam(pol.parameters(), lr=3e-4)
        algo = tg.GRPO(0.15, 0.5, 0.5, pol, opt, None, updates_per_iter=1)
        env_fn = lambda: tg.CartPole(max_steps=40)
        mgr = tg.RolloutManager(env_fn, pol, restart=False, num_workers=4, num_episodes_per_worker=5,
                                use_multiprocessing=False, seed=1)
        buf = tg.Rollout_Buffer(mgr)
        return tg.Pipeline("cartpole_nn_grpo", "001", env_fn, pol, algo, mgr, buf, None, None, load_path=load_path,
                           save_freq=1, root=str(tmp_path)), pol, opt, buf

    pipe, pol, opt, buf = build()
    pipe.train(3)
    d = pipe.archive_path
    assert d.endswith(os.path.join("archive", "CartPole", "cartpole_nn_grpo", "001"))
    for f in ("policy.pt", "optimizer.pth", "reward.csv", "metadata.json"):
        assert os.path.exists(os.path.join(d, f)), f
    meta = json.load(open(os.path.join(d, "metadata.json")))
    assert meta["env_name"] == "CartPole" and meta["policy"]["hidden_dims"] == [32, 32]
    assert meta["algorithm"] == {"algorithm": "GRPO", "epsilon": 0.15, "beta": 0.5, "updates_per_iter": 1}
    assert meta["policy"]["num_parameters"] == 5 * 32 + 32 + 32 * 32 + 32 + 32 + 1
    pipe2, pol2, opt2, buf2 = build(load_path=d)
    assert torch.equal(pol2.flat_parameters(), pol.flat_parameters())
    assert len(buf2.avg_reward) == 3 and np.allclose(buf2.avg_reward, [float(x) for x in buf.avg_reward])
    assert pipe2.loaded_metadata["checkpoint_name"] == "001"
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert s1.keys() == s2.keys()
    for k in s1:
        assert torch.equal(s1[k]["exp_avg"].cpu(), s2[k]["exp_avg"].cpu())
        assert float(s1[k]["step"]) == float(s2[k]["step"]) == 3.0
    pipe2.save_trajectory()
    assert os.path.exists(os.path.join(pipe2.archive_path, "trajectory.csv"))
    pipe2.publish()
    assert os.path.exists(os.path.join(pipe2.publish_path, "metadata.json"))


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
