"""Host-side mirror of models/neural_network.py and policies/actor_critic.py.

`NeuralNetwork` keeps the reference's module structure (a torch.nn.Sequential
of Linear + activation, state-dict keys `network.{0,2,...}.{weight,bias}`,
models/neural_network.py:36-65) so shipped checkpoints load unchanged, but its
parameters are views into ONE flat fp32 CUDA vector in torch order -- the layout
the C ABI consumes (rollout, gradient, Adam, NCCL allreduce all see one buffer).
The Gaussian policies evaluate through the kernels (tg_policy_forward); there is
no torch-autograd path and no CPU path.
"""
from __future__ import annotations

import math
import os
from abc import ABC, abstractmethod
from typing import Union

import numpy as np
import torch

from . import _lib as L
from . import engine


def _default_device():
    if not torch.cuda.is_available():
        raise L.EngineError("no CUDA device: the trajopt_grpo_b200 engine has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class NeuralNetwork(torch.nn.Module):
    """models/neural_network.py:4-77."""

    def __init__(self, input_dim: int, output_dim: int, hidden_dims: list, activation: Union[str, list] = "ReLU",
                 device=None):
        super().__init__()
        self.input_dim, self.output_dim, self.hidden_dims = input_dim, output_dim, list(hidden_dims)
        if hidden_dims:
            if isinstance(activation, str):
                activations = [activation] * len(hidden_dims)
            elif isinstance(activation, list):
                assert len(activation) == len(hidden_dims), \
                    "Number of activation functions must equal the number of hidden layers."
                activations = activation
            else:
                raise TypeError("activation must be either a string or a list of strings.")
            layers = []
            dims = [input_dim] + list(hidden_dims)
            for i in range(len(hidden_dims)):
                layers.append(torch.nn.Linear(dims[i], dims[i + 1]))
                layers.append(getattr(torch.nn, activations[i])())
            layers.append(torch.nn.Linear(hidden_dims[-1], output_dim))
            # what the kernels are told: one name, or the per-layer list (FP32-pipe kernels)
            self.activation_name = activations[0] if len(set(activations)) == 1 else list(activations)
        else:
            layers = [torch.nn.Linear(input_dim, output_dim)]
            activations = []
            self.activation_name = "ReLU"
        for nm in activations:
            if nm not in L.ACT_IDS:
                raise L.EngineError(f"activation {nm!r} is not supported by the kernels ({sorted(L.ACT_IDS)})")
        self.network = torch.nn.Sequential(*layers)
        self.dims = [input_dim] + list(hidden_dims) + [output_dim]
        self._flat = None
        self._device = torch.device(device) if device is not None else None

    # -- flat parameter vector ---------------------------------------------------
    def n_params(self) -> int:
        return sum(p.numel() for p in self.network.parameters())

    def bind_flat(self, flat: torch.Tensor):
        """Re-home every parameter as a view into `flat` (CUDA fp32, torch order)."""
        off = 0
        with torch.no_grad():
            for p in self.network.parameters():
                n = p.numel()
                view = flat[off:off + n].view(p.shape)
                view.copy_(p.detach().to(flat.device, torch.float32))
                p.data = view
                off += n
        self._flat = flat[:off]

    def flat_params(self) -> torch.Tensor:
        """The flat CUDA vector the kernels read; re-binds if a caller moved or
        replaced parameter storage (e.g. `.to()`)."""
        ok = self._flat is not None
        if ok:
            off = 0
            base = self._flat.data_ptr()
            for p in self.network.parameters():
                if p.data_ptr() != base + 4 * off or not p.is_cuda:
                    ok = False
                    break
                off += p.numel()
        if not ok:
            dev = self._device or _default_device()
            self.bind_flat(torch.empty(self.n_params(), dtype=torch.float32, device=dev))
        return self._flat

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """models/neural_network.py:67-77 through tg_policy_forward (rows = samples)."""
        flat = self.flat_params()
        x = torch.as_tensor(x, dtype=torch.float32, device=flat.device)
        lead = x.shape[:-1]
        xt = x.reshape(-1, self.input_dim).t().contiguous()
        mu, _ = engine.policy_forward(self.dims, self.activation_name, flat, xt)
        return mu.t().reshape(*lead, self.output_dim)


class ActorCritic(ABC):
    """policies/actor_critic.py:9-26."""

    @abstractmethod
    def forward(self, state):
        pass

    @abstractmethod
    def parameters(self):
        pass

    def __call__(self, state):
        return self.forward(state)


class RandomUniformActorCritic(ActorCritic):
    """policies/actor_critic.py:28-71 (no parameters; not usable by the fused rollout)."""

    def __init__(self, action_dim: int):
        self.action_dim = action_dim

    def forward(self, state):
        return torch.rand(self.action_dim) * 2 - 1, torch.zeros(1), torch.zeros(1)

    def parameters(self):
        return []


class _GaussianBase(ActorCritic):
    def _init_common(self, input_dim, output_dim, hidden_dims, activation, cov, device):
        self.input_dim, self.output_dim = input_dim, output_dim
        self.hidden_dims, self.activation = hidden_dims, activation
        if isinstance(cov, list):
            self.cov = torch.diag(torch.tensor(cov))               # actor_critic.py:100-103
        else:
            self.cov = torch.diag(torch.tensor([cov] * output_dim))
        self.device = torch.device(device) if device is not None else _default_device()

    @property
    def cov_diag(self):
        return [float(c) for c in torch.diagonal(self.cov)]

    def param_tag(self):
        """Identity of the current parameter VALUES without a device sync.  The Parameters are re-homed
        views of the flat buffer (`p.data = view`), so each has its OWN version counter: the sum of the
        parameters' `_version` sees every torch-side in-place write (load_state_dict, copy_, add_, a torch
        optimizer's step), `_param_epoch` counts the raw-pointer writes of tg_adam_step and every load /
        re-bind.  Writes that bypass both (`p.data.add_()`) are invisible to torch itself."""
        flat = self.flat_parameters()
        return (flat.data_ptr(), sum(p._version for p in self.parameters()), getattr(self, "_param_epoch", 0))

    def bump_param_epoch(self):
        self._param_epoch = getattr(self, "_param_epoch", 0) + 1

    def entropy_value(self) -> float:
        """MultivariateNormal.entropy() for the fixed diagonal covariance (constant)."""
        A = self.output_dim
        return 0.5 * A * (1.0 + math.log(2 * math.pi)) + sum(0.5 * math.log(c) for c in self.cov_diag)

    def _mean(self, obs_rows: torch.Tensor) -> torch.Tensor:
        xt = obs_rows.reshape(-1, self.input_dim).t().contiguous()
        mu, _ = engine.policy_forward(self.actor.dims, self.actor.activation_name, self.actor.flat_params(), xt)
        return mu.t()

    def forward(self, state):
        """actor_critic.py:107-138 / 255-289: one policy call for one observation.
        The noise comes from torch's global CPU generator, like the reference's
        MultivariateNormal.sample()."""
        st = torch.as_tensor(np.asarray(state), dtype=torch.float32).to(self.device)
        mu = self._mean(st.reshape(1, -1))[0]
        eps = torch.randn(self.output_dim).to(self.device)
        sd = torch.sqrt(torch.diagonal(self.cov)).to(self.device)
        action = mu + sd * eps
        z = (action - mu) / sd
        log_prob = -0.5 * (self.output_dim * math.log(2 * math.pi) + (z * z).sum()) - torch.log(sd).sum()
        return action.cpu().numpy(), log_prob, self._value_one(st)

    def log_prob(self, observation, action):
        """actor_critic.py:140-160 / 291-311 -> (log_prob [n], entropy [n])."""
        flat = self.actor.flat_params()
        obs = torch.as_tensor(np.asarray(observation) if not torch.is_tensor(observation) else observation,
                              dtype=torch.float32).to(flat.device)
        act = torch.as_tensor(np.asarray(action) if not torch.is_tensor(action) else action,
                              dtype=torch.float32).to(flat.device)
        lead = obs.shape[:-1]
        xt = obs.reshape(-1, self.input_dim).t().contiguous()
        at = act.reshape(-1, self.output_dim).t().contiguous()
        _, lp = engine.policy_forward(self.actor.dims, self.actor.activation_name, flat, xt, self.cov_diag, at,
                                      want_mu=False, want_logp=True)
        ent = torch.full_like(lp, self.entropy_value())
        return lp.reshape(lead), ent.reshape(lead)

    def metadata(self):
        return {
            "input_dim": self.input_dim,
            "output_dim": self.output_dim,
            "hidden_dims": self.hidden_dims,
            "activation": self.activation,
            "cov": self.cov.tolist() if isinstance(self.cov, torch.Tensor) else self.cov,
            "num_parameters": sum(p.numel() for p in self.parameters()),
        }


class GaussianActor_NeuralNetwork(_GaussianBase):
    """policies/actor_critic.py:73-215 (actor only, fixed diagonal covariance)."""

    def __init__(self, input_dim: int, output_dim: int, hidden_dims: Union[list, tuple], activation: str = "ReLU",
                 cov: Union[list, float] = 0.1, device=None):
        self._init_common(input_dim, output_dim, hidden_dims, activation, cov, device)
        self.actor = NeuralNetwork(input_dim, output_dim, list(hidden_dims), activation, device=self.device)
        self._flat_all = torch.empty(self.actor.n_params(), dtype=torch.float32, device=self.device)
        self.actor.bind_flat(self._flat_all)

    def _value_one(self, st):
        return None

    def value(self, state):
        return [None] * state.shape[0]                             # actor_critic.py:162-172

    def parameters(self):
        return self.actor.parameters()

    def state_dict(self):
        return self.actor.state_dict()

    def load_state_dict(self, state_dict):
        self.actor.load_state_dict(state_dict)
        self.bump_param_epoch()

    def flat_parameters(self) -> torch.Tensor:
        self.actor.flat_params()
        return self.actor._flat

    def save(self, path):
        if L.is_checkpoint_writer():
            torch.save({k: v.cpu() for k, v in self.actor.state_dict().items()}, os.path.join(path, "policy.pt"))

    def load(self, path):
        """Missing in the reference (GRPO resume raises); added so
        checkpoints written by save() round-trip.  `checkpoint_loads` lets GRPO re-synchronise its
        old_policy with the restored weights (a checkpoint is written right after a learn(), where
        old_policy == policy, grpo.py:148)."""
        self.load_state_dict(torch.load(os.path.join(path, "policy.pt"), weights_only=True))
        self.checkpoint_loads = getattr(self, "checkpoint_loads", 0) + 1


class GaussianActorCritic_NeuralNetwork(_GaussianBase):
    """policies/actor_critic.py:220-378 (actor + critic MLP of the same hidden shape)."""

    def __init__(self, input_dim: int, output_dim: int, hidden_dims: Union[list, tuple], activation: str = "ReLU",
                 cov: Union[list, float] = 0.1, device=None):
        self._init_common(input_dim, output_dim, hidden_dims, activation, cov, device)
        self.actor = NeuralNetwork(input_dim, output_dim, list(hidden_dims), activation, device=self.device)
        self.critic = NeuralNetwork(input_dim, 1, list(hidden_dims), activation, device=self.device)
        na, nc = self.actor.n_params(), self.critic.n_params()
        self._flat_all = torch.empty(na + nc, dtype=torch.float32, device=self.device)
        self.actor.bind_flat(self._flat_all[:na])
        self.critic.bind_flat(self._flat_all[na:])

    def _value_one(self, st):
        return self.critic(st.reshape(1, -1))[0]

    def value(self, state):
        """actor_critic.py:313-323."""
        st = torch.as_tensor(np.asarray(state) if not torch.is_tensor(state) else state, dtype=torch.float32)
        return self.critic(st.to(self.device)).squeeze()

    def parameters(self):
        return list(self.actor.parameters()) + list(self.critic.parameters())

    def state_dict(self):
        return {"actor": self.actor.state_dict(), "critic": self.critic.state_dict()}

    def load_state_dict(self, state_dict):
        self.actor.load_state_dict(state_dict["actor"])
        self.critic.load_state_dict(state_dict["critic"])
        self.bump_param_epoch()

    def load(self, path):
        """actor_critic.py:340-346."""
        sd = torch.load(os.path.join(path, "policy.pt"), weights_only=True)
        self.load_state_dict(sd)

    def flat_parameters(self) -> torch.Tensor:
        """actor parameters followed by critic parameters, one buffer."""
        a, c = self.actor.flat_params(), self.critic.flat_params()
        if a.data_ptr() + 4 * a.numel() != c.data_ptr():
            flat = torch.empty(a.numel() + c.numel(), dtype=torch.float32, device=self.device)
            self.actor.bind_flat(flat[:a.numel()])
            self.critic.bind_flat(flat[a.numel():])
            self._flat_all = flat
        return self._flat_all

    def save(self, save_path):
        if L.is_checkpoint_writer():
            sd = {k: {kk: vv.cpu() for kk, vv in v.items()} for k, v in self.state_dict().items()}
            torch.save(sd, os.path.join(save_path, "policy.pt"))
