"""CPU checks of the benchmark's start policies (bench_assets/*.json, designed by oracle/make_start_policy.py):
the linear feedback is embedded EXACTLY in the ReLU network bench.py builds, and under the full exploration noise it
keeps the oracle's quadrotor envs alive (a random-init policy crashes within ~70 steps)."""
import json
import os
import sys

import numpy as np
import pytest

import restate as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("workload,asset,eq_ones", [("quadpole_cfg4", "quadpole_lqr_gain.json", (6, 13)),
                                                    ("quadpole2d_cfg3", "quadpole2d_lqr_gain.json", (5, 8))])
def test_start_policy_is_the_embedded_linear_feedback_and_keeps_the_envs_alive(workload, asset, eq_ones):
    w = bench.WORKLOADS[workload]
    kind, T = w["kind"], 300
    g = json.load(open(os.path.join(ROOT, "bench_assets", asset)))
    K, sel = np.asarray(g["K"]), g["sel"]
    dims, Ws, bs = bench.start_policy_arrays(w)
    assert dims == [R.OBS_DIM[kind]] + w["hidden"] + [R.ACT_DIM[kind]]
    rng = np.random.default_rng(0)
    x = rng.standard_normal((64, dims[0]))
    eq = np.zeros(dims[0])
    for i in eq_ones:
        eq[i] = 1.0
    mu = R.mlp_forward(x, Ws, bs, R.ACT_RELU, np.float64)
    np.testing.assert_allclose(mu, -(x - eq)[:, sel] @ K.astype(np.float32).T, rtol=1e-6, atol=1e-6)
    # closed loop in the oracle under the workload's exploration noise
    N = 128
    s = R.reset_states(kind, N, rng)
    alive = np.ones(N, bool)
    sd = np.float32(np.sqrt(w["cov"]))
    step = R.quadpole_step if kind == R.ENV_QUADPOLE else R.quadpole2d_step
    for _ in range(T):
        a = R.mlp_forward(s, Ws, bs, R.ACT_RELU, np.float32) + sd * rng.standard_normal((N, dims[-1])).astype(np.float32)
        nxt, _, aux = step(s, a.astype(np.float32), 0.02, np.float64)
        alive &= ~aux["oob"]
        s = np.where(alive[:, None], nxt, s)
    assert alive.mean() >= 0.99, alive.mean()
    assert g["alive_fraction_oracle"] >= 0.99 and g["valid_fraction_zero_policy"] < 0.2
