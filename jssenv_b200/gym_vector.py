"""Gymnasium ``VectorEnv``-style adapter over ``JssVecEnv``.

gymnasium is optional (it is not installed in the build image): when it is importable the
adapter subclasses ``gymnasium.vector.VectorEnv`` and publishes batched spaces; otherwise it is
a plain class with the same methods.  Semantics follow gymnasium's "next-step" autoreset mode,
which is exactly what ``JssVecEnv(auto_reset=True)`` implements on device: the step after a
terminal one performs the reset (reward 0, terminated False).  Observations stay on the GPU
(torch tensors) unless ``to_numpy=True``.
"""
from typing import Any, Dict, Optional

import numpy as np

from .vec_env import JssVecEnv

try:
    import gymnasium as _gym
    _Base = _gym.vector.VectorEnv
except Exception:  # pragma: no cover - gymnasium absent
    _gym = None
    _Base = object


class JssGymVectorEnv(_Base):
    metadata = {"autoreset_mode": "next_step"}

    def __init__(self, num_envs: int, env_config: Optional[Dict[str, Any]] = None, device: int = 0,
                 to_numpy: bool = False, seed: int = 0):
        self.vec = JssVecEnv(num_envs, env_config, device=device, auto_reset=True, seed=seed)
        self.num_envs = num_envs
        self.to_numpy = to_numpy
        J = self.vec.jobs
        if _gym is not None:
            single_obs = _gym.spaces.Dict({
                "action_mask": _gym.spaces.Box(0, 1, shape=(J + 1,), dtype=np.int8),
                "real_obs": _gym.spaces.Box(0.0, 1.0, shape=(J, 7), dtype=np.float32)})
            self.single_observation_space = single_obs
            self.single_action_space = _gym.spaces.Discrete(J + 1)
            self.observation_space = _gym.vector.utils.batch_space(single_obs, num_envs)
            self.action_space = _gym.vector.utils.batch_space(self.single_action_space, num_envs)

    def _out(self, obs):
        if not self.to_numpy:
            return obs
        return {k: v.cpu().numpy() for k, v in obs.items()}

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.vec.seed = int(seed)
        return self._out(self.vec.reset()), {}

    def step(self, actions):
        obs, reward, done, truncated, info = self.vec.step(actions)
        if self.to_numpy:
            return self._out(obs), reward.cpu().numpy(), done.cpu().numpy(), truncated.cpu().numpy(), info
        return obs, reward, done, truncated, info

    def close(self, **kwargs):
        self.vec.close()
