"""Design the stabilising START POLICY of the quadrotor benchmark workloads (bench.py).

TEST / BENCH INFRASTRUCTURE -- runs in the build container (it uses the oracle's dynamics,
oracle/restate.py: quadpole_step, for the linearisation and for the survival check); bench.py only reads
the small JSON it writes (bench_assets/quadpole_lqr_gain.json), never this module.

Why: the quadrotor envs end an episode when the vehicle leaves its +-1.5 m box
(quadrotor_env.py:699-708).  A freshly initialised policy with the reference's exploration noise (cov 0.3,
quadpole_pipeline_ppo.py:58) crashes within ~70 steps (valid fraction 0.07 of N*T), so a benchmark started from
random weights measures the zero-fill path.  The benchmark therefore starts from a policy a few hundred
training epochs would reach: a linear state feedback a = -K z that holds the vehicle near the origin and damps
the payload swing UNDER the full cov-0.3 exploration noise, embedded exactly in the 20-256-256-4 ReLU network
(z = relu(z) - relu(-z)); every other weight keeps torch's default initialisation.

K is the discrete LQR gain of the dynamics linearised (central differences through the oracle step) about
hover, on the 16 coordinates that are zero at equilibrium (positions, velocities, the vector parts of both
quaternions, body rates, the two swing rates).

    python oracle/make_start_policy.py        # writes bench_assets/quadpole_lqr_gain.json
"""
import json
import os
import sys

import numpy as np
import scipy.linalg as sla

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import restate as R  # noqa: E402

DT = 0.02
EQ = np.zeros(20); EQ[6] = 1.0; EQ[13] = 1.0                       # hover: identity attitude, payload straight down
SEL = [0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 11, 12, 14, 15, 17, 18]      # coordinates that vanish at equilibrium


def _f(z, a):
    s = np.tile(EQ, (z.shape[0], 1))
    s[:, SEL] += z
    nxt, _, _ = R.quadpole_step(s, a.astype(np.float32), DT, np.float64)
    return (nxt - EQ)[:, SEL]


def linearise():
    nz = len(SEL)
    A, B = np.zeros((nz, nz)), np.zeros((nz, 4))
    z0, a0 = np.zeros((1, nz)), np.zeros((1, 4))
    for i in range(nz):
        e = np.zeros((1, nz)); e[0, i] = 1e-4
        A[:, i] = (_f(e, a0) - _f(-e, a0))[0] / 2e-4
    for j in range(4):
        e = np.zeros((1, 4)); e[0, j] = 1e-2                       # actions are float32
        B[:, j] = (_f(z0, e) - _f(z0, -e))[0] / 2e-2
    return A, B


def lqr(A, B):
    Q = np.diag([10.0] * 3 + [1.0] * 3 + [10.0] * 3 + [1.0] * 3 + [1.0] * 2 + [1.0] * 2)
    Rm = np.eye(4)
    P = sla.solve_discrete_are(A, B, Q, Rm)
    return np.linalg.solve(Rm + B.T @ P @ B, B.T @ P @ A)


def survival(K, N=2048, T=1000, cov=0.3, seed=0):
    """Fraction of envs alive after T steps and valid fraction of N*T under a = -K z + sqrt(cov) eps."""
    rng = np.random.default_rng(seed)
    s = R.reset_states(R.ENV_QUADPOLE, N, rng)
    alive, length = np.ones(N, bool), np.zeros(N, int)
    sd = np.float32(np.sqrt(cov))
    for _ in range(T):
        mu = (-((s - EQ)[:, SEL] @ K.T)).astype(np.float32)
        a = mu + sd * rng.standard_normal((N, 4)).astype(np.float32)
        nxt, _, aux = R.quadpole_step(s, a, DT, np.float64)
        length += alive
        alive &= ~aux["oob"]
        s = np.where(alive[:, None], nxt, s)
    return float(alive.mean()), float(length.mean() / T)


# ---- QuadPole2D (BASELINE configs[2]): state feedback on the vehicle only (x, z, vx, vz, sin(theta), theta_dot).
# The pole starts anywhere on the circle (quadrotor_env.py:951-955) and has no damping, so it keeps swinging; it is
# treated as a disturbance of the vehicle, which only has to stay inside its +-2 m box (quadrotor_env.py:1020-1022).
EQ2 = np.zeros(10); EQ2[5] = 1.0; EQ2[8] = 1.0
SEL2 = [0, 1, 2, 3, 4, 6]


def design_2d():
    def f(z, a):
        s = np.tile(EQ2, (z.shape[0], 1))
        s[:, SEL2] += z
        nxt, _, _ = R.quadpole2d_step(s, a.astype(np.float32), DT, np.float64)
        return (nxt - EQ2)[:, SEL2]
    n = len(SEL2)
    A, B = np.zeros((n, n)), np.zeros((n, 2))
    z0, a0 = np.zeros((1, n)), np.zeros((1, 2))
    for i in range(n):
        e = np.zeros((1, n)); e[0, i] = 1e-4
        A[:, i] = (f(e, a0) - f(-e, a0))[0] / 2e-4
    for j in range(2):
        e = np.zeros((1, 2)); e[0, j] = 1e-2
        B[:, j] = (f(z0, e) - f(z0, -e))[0] / 2e-2
    Q, Rm = np.diag([10.0, 10.0, 1.0, 1.0, 10.0, 1.0]), np.eye(2)
    P = sla.solve_discrete_are(A, B, Q, Rm)
    return np.linalg.solve(Rm + B.T @ P @ B, B.T @ P @ A)


def survival_2d(K, N=2048, T=500, cov=0.5, seed=0):
    rng = np.random.default_rng(seed)
    s = R.reset_states(R.ENV_QUADPOLE2D, N, rng)
    alive, length = np.ones(N, bool), np.zeros(N, int)
    sd = np.float32(np.sqrt(cov))
    for _ in range(T):
        mu = (-((s - EQ2)[:, SEL2] @ K.T)).astype(np.float32)
        a = mu + sd * rng.standard_normal((N, 2)).astype(np.float32)
        nxt, _, aux = R.quadpole2d_step(s, a, DT, np.float64)
        length += alive
        alive &= ~aux["oob"]
        s = np.where(alive[:, None], nxt, s)
    return float(alive.mean()), float(length.mean() / T)


def main():
    K2 = design_2d()
    alive2, valid2 = survival_2d(K2)
    _, valid20 = survival_2d(np.zeros_like(K2), N=512)
    out2 = {"env": "QuadPole2D", "sel": SEL2, "K": K2.tolist(), "cov": 0.5, "horizon": 500,
            "alive_fraction_oracle": alive2, "valid_fraction_oracle": valid2, "valid_fraction_zero_policy": valid20,
            "how": "discrete LQR (Q = diag(10 pos, 1 vel, 10 sin(theta), 1 theta_dot), R = I) of the oracle's quadpole2d_step "
                   "linearised about hover on the vehicle coordinates only; survival measured with oracle/restate.py on 2048 envs"}
    path2 = os.path.join(os.path.dirname(HERE), "bench_assets", "quadpole2d_lqr_gain.json")
    with open(path2, "w") as f:
        json.dump(out2, f, indent=1)
    print(f"QuadPole2D: alive after 500 steps {alive2:.4f}, valid fraction {valid2:.4f} (zero policy: {valid20:.3f}) -> {path2}")
    A, B = linearise()
    K = lqr(A, B)
    alive, valid = survival(K)
    _, valid0 = survival(np.zeros_like(K), N=512)
    out = {"env": "QuadPole", "sel": SEL, "K": K.tolist(), "cov": 0.3, "horizon": 1000,
           "alive_fraction_oracle": alive, "valid_fraction_oracle": valid, "valid_fraction_zero_policy": valid0,
           "how": "discrete LQR (Q = diag(10 pos, 1 vel, 10 att, 1 rate, 1 swing, 1 swing rate), R = I) of the oracle's "
                  "quadpole_step linearised about hover; survival measured with oracle/restate.py on 2048 envs"}
    path = os.path.join(os.path.dirname(HERE), "bench_assets", "quadpole_lqr_gain.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(f"alive after 1000 steps {alive:.4f}, valid fraction {valid:.4f} (zero policy: {valid0:.3f}) -> {path}")


if __name__ == "__main__":
    main()
