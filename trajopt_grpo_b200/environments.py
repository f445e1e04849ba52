"""Host-side mirror of the reference's environment classes.

Same names, constructor arguments, attributes and `reset / restart / step`
protocol as environments/env.py:10-68 and the concrete classes (CartPole
cartpole_env.py:6-182, Pendulum pendulum_env.py:7-162, QuadPole
quadrotor_env.py:353-713, QuadPole2D quadrotor_env.py:867-1223, Quadrotor
quadrotor_env.py:6-182).  These objects are DESCRIPTORS of an environment for
the fused rollout kernel (kind id, timestep, horizon, reset distribution); a
single `step()` call goes through the batched `tg_env_step` kernel with N = 1 in
float64 -- there is no host arithmetic path.  Rendering is out of scope.
"""
from __future__ import annotations

import abc

import numpy as np
import torch

from . import _lib as L
from . import engine


class Box:
    """Minimal stand-in for gymnasium.spaces.Box (only `.shape`, `.low`, `.high`,
    `.dtype`, `sample`, `contains` are used by the reference and its tests)."""

    def __init__(self, low, high, shape, dtype=np.float32):
        self.low = np.full(shape, low, dtype)
        self.high = np.full(shape, high, dtype)
        self.shape = tuple(shape)
        self.dtype = dtype

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return np.random.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(((x >= self.low) & (x <= self.high)).all())


class Env(abc.ABC):
    """environments/env.py:10-68 (gym.Env + `restart`)."""

    _tg_kind: int = -1
    _tg_phys = None          # physical constructor arguments for the kernels (None = reference defaults)
    _state_keys: tuple = ()
    _state_split: tuple = ()

    def __init__(self, env_name: str) -> None:
        self.env_name = env_name

    # -- reset distribution, vectorised (used by RolloutManager) ---------------
    @abc.abstractmethod
    def sample_initial_states(self, n: int, rng: np.random.Generator) -> np.ndarray:
        """[n, S] float64 draws from this env's reset() distribution."""

    # -- gymnasium-style protocol ----------------------------------------------
    def _set_full_state(self, vec):
        off = 0
        for key, k in zip(self._state_keys, self._state_split):
            self.state_dict[key] = np.array(vec[off:off + k], dtype=np.float64)
            off += k

    def reset(self):
        vec = self.sample_initial_states(1, _GLOBAL_RNG())[0]
        self._set_full_state(vec)
        self._initial_state = {k: v.copy() for k, v in self.state_dict.items()}
        self._steps = 0
        self._time = 0
        self._time_balanced = 0
        self._bal_count = 0
        return self._get_obs(), self._get_info()

    def restart(self):
        self.state_dict = {k: v.copy() for k, v in self._initial_state.items()}
        self._steps = 0
        self._time = 0
        self._time_balanced = 0
        self._bal_count = 0
        return self._get_obs(), self._get_info()

    def _get_obs(self):
        return np.hstack([self.state_dict[k] for k in self._state_keys])

    def _get_info(self):
        return {"time_balanced": self._time_balanced}

    def _step_device(self, action):
        s = torch.from_numpy(self._get_obs().reshape(-1, 1).copy()).cuda()
        a = torch.as_tensor(np.asarray(action, dtype=np.float32).reshape(-1, 1)).cuda()
        steps = torch.tensor([self._steps], dtype=torch.int32, device="cuda")
        bal = torch.tensor([self._bal_count], dtype=torch.int32, device="cuda")
        nxt, rew, done, bal_out = engine.env_step(self._tg_kind, self.max_steps, self.timestep, s, a, steps, bal,
                                                  phys=self._tg_phys)
        return nxt.cpu().numpy()[:, 0], float(rew.item()), bool(done.item()), int(bal_out.item())

    def step(self, action):
        nxt, reward, done, bal = self._step_device(action)
        self._set_full_state(nxt)
        self._steps += 1
        self._time += self.timestep
        self._bal_count = bal
        self._time_balanced = bal * self.timestep
        terminated, truncated = self._split_done(done)
        return self._get_obs(), reward, terminated, truncated, self._get_info()

    def _split_done(self, done):
        return False, done            # the quadrotor envs and CartPole only ever truncate

    def _dynamics(self, state, control):
        """Env._dynamics(state, control) (cartpole_env.py:51-92, pendulum_env.py:48-75, quadrotor_env.py:417-528,
        1044-1130): one transition from `state` under the ALREADY WRAPPED `control` (float32, as step() passes
        it), through tg_env_dynamics in float64."""
        s = torch.from_numpy(np.asarray(state, np.float64).reshape(-1, 1).copy()).cuda()
        u = torch.from_numpy(np.asarray(control, np.float32).reshape(-1, 1).copy()).cuda()
        return engine.env_dynamics(self._tg_kind, self.timestep, s, u, phys=self._tg_phys).cpu().numpy()[:, 0]

    def render(self, *a, **k):
        raise NotImplementedError("rendering is outside the hot path (SURVEY.md section 2)")


def _GLOBAL_RNG():
    # the reference draws from numpy's global RNG in reset(); keep that stream
    class _G:
        @staticmethod
        def uniform(lo, hi, n):
            return np.random.uniform(lo, hi, n)
    return _G


class CartPole(Env):
    """cartpole_env.py:6-182."""
    _tg_kind = L.ENV_CARTPOLE
    _state_keys = ("cartpole",)
    _state_split = (5,)

    def __init__(self, env_name: str = "CartPole", masscart: float = 1.0, masspole: float = 1.0, length: float = 0.5,
                 gravity: float = 9.80665, timestep: float = 0.02, max_steps: int = 500):
        super().__init__(env_name)
        self.masscart, self.masspole, self.length, self.gravity = masscart, masspole, length, gravity
        self._tg_phys = (masscart, masspole, length, gravity)
        self.timestep, self.max_steps = timestep, max_steps
        self.max_time = max_steps * timestep
        self._initial_state = None
        self._steps = self._time = self._time_balanced = self._bal_count = 0
        self.state_dict = {"cartpole": np.zeros(5)}
        self._is_3d = False
        self.observation_space = Box(-1, 1, (5,), np.float32)
        self.action_space = Box(-1, 1, (1,), np.float32)

    def _wrap_action(self, action):
        return 5 * np.clip(action, -1, 1)

    def sample_initial_states(self, n, rng):
        th = rng.uniform(-np.pi, np.pi, n)          # cartpole_env.py:103
        z = np.zeros(n)
        return np.stack([z, z, np.sin(th), np.cos(th), z], 1)


class Pendulum(Env):
    """pendulum_env.py:7-162."""
    _tg_kind = L.ENV_PENDULUM
    _state_keys = ("pendulum",)
    _state_split = (3,)

    def __init__(self, env_name: str = "Pendulum", swingup: bool = False, mass: float = 1.0, length: float = 0.5,
                 gravity: float = 9.80665, timestep: float = 0.05, max_steps: int = 200):
        super().__init__(env_name)
        self.swingup, self.mass, self.length, self.gravity = swingup, mass, length, gravity
        self._tg_phys = (mass, length, gravity)
        self.timestep, self.max_steps = timestep, max_steps
        self.max_time = max_steps * timestep
        self._initial_state = None
        self._steps = self._time = self._time_balanced = self._bal_count = 0
        self.state_dict = {"pendulum": np.zeros(3)}
        self._is_3d = False
        self.observation_space = Box(-1, 1, (3,), np.float32)
        self.action_space = Box(-1, 1, (1,), np.float32)

    def _wrap_action(self, action):
        return np.clip(action, -1, 1)

    def sample_initial_states(self, n, rng):
        if self.swingup:
            th = rng.uniform(-np.pi, np.pi, n)      # pendulum_env.py:88-91
        else:
            th = rng.uniform(np.pi - 0.05, np.pi + 0.05, n)
        return np.stack([np.sin(th), np.cos(th), np.zeros(n)], 1)

    def step(self, action):
        # pendulum_env.py:162 returns (obs, reward, truncated, terminated, info) -- swapped
        nxt, reward, done, bal = self._step_device(action)
        self._set_full_state(nxt)
        self._steps += 1
        self._time += self.timestep
        self._bal_count = bal
        self._time_balanced = bal * self.timestep
        terminated = bal >= L.balanced_limit_count(self.timestep)
        truncated = self._steps >= L.time_limit_step(self.timestep, self.max_steps)
        return self._get_obs(), reward, truncated, terminated, self._get_info()


class QuadPole2D(Env):
    """quadrotor_env.py:867-1223."""
    _tg_kind = L.ENV_QUADPOLE2D
    _state_keys = ("quadrotor", "pendulum")
    _state_split = (7, 3)

    def __init__(self, env_name="QuadPole2D", max_steps=500, timestep=0.02):
        super().__init__(env_name)
        self.mq, self.mp, self.I, self.Lq, self.Lp = 1.5, 0.5, 4e-1, 0.5, 0.75
        self.gravity, self.timestep, self.max_steps = 9.80665, timestep, max_steps
        self.spatial_bounds = ((-2.0, 2.0), (-2.0, 2.0))
        self.balance_radius = 0.25
        self._is_3d = False
        self._xbounds, self._zbounds = self.spatial_bounds
        self.hover_force = (self.mq + self.mp) * self.gravity / 2
        self.state_dict = {"quadrotor": np.zeros(7), "pendulum": np.zeros(3)}
        self._initial_state = None
        self._steps = self._time = self._time_balanced = self._bal_count = 0
        self.observation_space = Box(-np.inf, np.inf, (10,), np.float32)
        self.action_space = Box(0.0, 20.0, (2,), np.float32)

    def _wrap_action(self, action):
        return self.hover_force + self.hover_force * np.clip(action, -1, 1)

    def sample_initial_states(self, n, rng):
        ph = rng.uniform(-np.pi, np.pi, n)          # quadrotor_env.py:951-955
        z, o = np.zeros(n), np.ones(n)
        return np.stack([z, z, z, z, z, o, z, np.sin(ph), np.cos(ph), z], 1)


class QuadPole(Env):
    """quadrotor_env.py:353-713 (3-D quadrotor with a slung payload, quaternions)."""
    _tg_kind = L.ENV_QUADPOLE
    _state_keys = ("quadrotor", "pendulum")
    _state_split = (13, 7)

    def __init__(self, env_name="QuadPole", max_steps=500):
        super().__init__(env_name)
        self.max_steps = max_steps
        self.mass, self.load_mass, self.gravity, self.tether_length = 1.5, 0.5, 9.80665, 0.5
        self.Ixx, self.Iyy, self.Izz = 4e-1, 4e-1, 2.5e-1
        self.torque_constant, self.arm_length, self.timestep = 0.1, 0.5, 0.02
        self.hover_force = (self.mass + self.load_mass) * self.gravity / 4
        self.spatial_bounds = ((-1.5, 1.5), (-1.5, 1.5), (-1.5, 1.5))
        self._xbounds, self._ybounds, self._zbounds = self.spatial_bounds
        self.state_dict = {"quadrotor": np.zeros(13), "pendulum": np.zeros(7)}
        self._initial_state = None
        self._steps = self._time = self._time_balanced = self._bal_count = 0
        self._is_3d = True
        self.detailed_rendering = False
        self.observation_space = Box(-np.inf, np.inf, (20,), np.float32)
        self.action_space = Box(0.0, 20.0, (4,), np.float32)

    def _wrap_action(self, action):
        return self.hover_force + self.hover_force * np.clip(action, -1, 1)

    def sample_initial_states(self, n, rng):
        al = rng.uniform(-1.0, 1.0, n)              # quadrotor_env.py:543-544
        be = rng.uniform(-1.0, 1.0, n)
        z, o = np.zeros(n), np.ones(n)
        # q_p = normalize(q_y(beta) (x) q_x(alpha))  (:557-560)
        ca, sa, cb, sb = np.cos(al / 2), np.sin(al / 2), np.cos(be / 2), np.sin(be / 2)
        qp = np.stack([cb * ca, cb * sa, sb * ca, -sb * sa], 1)
        qp = qp / np.linalg.norm(qp, axis=1, keepdims=True)
        quad = np.stack([z, z, z, z, z, z, o, z, z, z, z, z, z], 1)
        return np.concatenate([quad, qp, np.zeros((n, 3))], 1)


class Quadrotor:
    """quadrotor_env.py:6-182: only `_dynamics` is functional in the reference
    (reset/step delegate to a non-existent self.env)."""

    def __init__(self, mass: float = 1.0, arm_length: float = 0.2, Ixx: float = 0.005, Iyy: float = 0.005,
                 Izz: float = 0.006, torque_constant: float = 0.017, gravity: float = 9.80665,
                 timestep: float = 0.05, max_steps: int = 200, **kw):
        if (mass, arm_length, Ixx, Iyy, Izz, torque_constant, gravity) != (1.0, 0.2, 0.005, 0.005, 0.006, 0.017, 9.80665):
            raise L.EngineError("the Quadrotor kernel is specialised for the reference's default parameters")
        self.timestep, self.max_steps = timestep, max_steps
        self.max_time = max_steps * timestep

    def _dynamics(self, state, control):
        s = torch.from_numpy(np.asarray(state, np.float64).reshape(12, 1).copy()).cuda()
        u = torch.from_numpy(np.asarray(control, np.float64).reshape(4, 1).copy()).cuda()
        return engine.quadrotor12_dynamics(s, u, self.timestep).cpu().numpy()[:, 0]
