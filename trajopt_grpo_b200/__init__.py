"""B200-native rollout-and-update engine behind the API of Dyllon-Preston/trajopt-grpo.

The reference's packages map onto this one as

    environments.{CartPole, Pendulum, QuadPole2D, QuadPole, Quadrotor, Env}  -> .environments
    rollout.{RolloutManager, RolloutWorker}                                   -> .rollout
    buffers.{Rollout_Buffer, TokenizedBuffer, Buffer}                         -> .buffers
    models.NeuralNetwork, policies.GaussianActor(Critic)_NeuralNetwork       -> .policies
    algorithms.{Algorithm, GRPO, PPO}                                         -> .algorithms

The reference's own orchestration (`pipelines/pipeline.py`: `buffer.sample()`, `algorithm.learn(buffer)`,
`{algorithm,policy,buffer}.{save,load,metadata}`, `rollout_manager.shutdown()`) drives these classes unchanged;
it is not re-implemented here.

with the arithmetic done by hand-written sm_100a kernels behind the C ABI in
include/trajopt_grpo.h (bound in ._lib / .engine).  There is no CPU fallback.
"""
from . import _lib
from .algorithms import GRPO, PPO, Algorithm
from .buffers import Buffer, Rollout_Buffer, TokenizedBuffer
from .environments import CartPole, Env, Pendulum, QuadPole, QuadPole2D, Quadrotor
from .policies import (ActorCritic, GaussianActor_NeuralNetwork, GaussianActorCritic_NeuralNetwork, NeuralNetwork,
                       RandomUniformActorCritic)
from .rollout import DeviceRollout, RolloutManager, RolloutWorker

__all__ = [
    "Algorithm", "GRPO", "PPO", "Buffer", "Rollout_Buffer", "TokenizedBuffer", "Env", "CartPole", "Pendulum",
    "QuadPole2D", "QuadPole", "Quadrotor", "ActorCritic", "RandomUniformActorCritic", "NeuralNetwork",
    "GaussianActor_NeuralNetwork", "GaussianActorCritic_NeuralNetwork", "RolloutManager", "RolloutWorker",
    "DeviceRollout",
]
