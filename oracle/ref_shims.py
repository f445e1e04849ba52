"""Import shims for running the UNMODIFIED reference in the build container.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs may touch it (and those never touch THIS file at run time on the
GPU box: /root/reference does not exist there).

The reference (Dyllon-Preston/trajopt-grpo, mounted read-only at
/root/reference) imports ``gymnasium`` and ``matplotlib`` at module top
(environments/env.py:1, environments/cartpole_env.py:3-4,
environments/quadrotor_env.py:4,232, buffers/rollout_buffer.py:3-4); neither is
installed here and only ``gym.Env`` (as a base class) and ``gym.spaces.Box``
(``.shape``) are used on the hot path.  ``install()`` puts tiny stand-ins into
``sys.modules`` and the reference root on ``sys.path`` so that
``make_golden.py`` can drive the real code to produce the committed fixtures.
"""
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("TRAJOPT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "environments"))


def install() -> None:
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")
        spaces = types.ModuleType("gymnasium.spaces")

        class Env:  # gym.Env stand-in; only __init__ is reached (environments/env.py:15)
            def __init__(self, *a, **k):
                pass

        class Box:  # only .shape is read by rollout code (rollout_manager.py:40-41)
            def __init__(self, low, high, shape, dtype=np.float32):
                self.low = np.full(shape, low, dtype)
                self.high = np.full(shape, high, dtype)
                self.shape = tuple(shape)
                self.dtype = dtype

        gym.Env, gym.spaces, spaces.Box = Env, spaces, Box
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces

    class _Any(types.ModuleType):  # attribute-permissive dummy (rendering is never called)
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return type(k, (object,), {"__init__": lambda s, *a, **kw: None})

    for n in ["matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.ticker",
              "matplotlib.gridspec", "mpl_toolkits", "mpl_toolkits.mplot3d",
              "mpl_toolkits.mplot3d.art3d"]:
        if n not in sys.modules:
            sys.modules[n] = _Any(n)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
