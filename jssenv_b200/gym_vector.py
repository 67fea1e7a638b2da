"""Gymnasium ``VectorEnv``-style adapter over ``JssVecEnv`` and the
``create_env`` helper mirroring ``JSSEnv/utils.py:32-60`` (the RLlib convenience of the reference).

gymnasium is optional (it is not installed in the build image): when it is importable the
adapter subclasses ``gymnasium.vector.VectorEnv`` and publishes batched spaces; otherwise it is
a plain class with the same methods.  Semantics follow gymnasium's "next-step" autoreset mode,
which is exactly what ``JssVecEnv(auto_reset=True)`` implements on device: the step after a
terminal one performs the reset (reward 0, terminated False).  Observations stay on the GPU
(torch tensors) unless ``to_numpy=True``.  ``action_mask`` is published as int8 0/1 (a zero-copy
view of the device bytes) so that data and the declared ``Box(0, 1, int8)`` space agree.
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


def _autoreset_mode():
    mode = getattr(getattr(_gym, "vector", None), "AutoresetMode", None) if _gym is not None else None
    return mode.NEXT_STEP if mode is not None else "next_step"      # enum since gymnasium 1.0


class JssGymVectorEnv(_Base):
    metadata = {"autoreset_mode": _autoreset_mode()}

    def __init__(self, num_envs: int, env_config: Optional[Dict[str, Any]] = None, device: int = 0,
                 to_numpy: bool = False, seed: int = 0):
        self.vec = JssVecEnv(num_envs, env_config, device=device, auto_reset=True, seed=seed)
        self.num_envs = num_envs
        self.to_numpy = to_numpy
        J = self.vec.jobs
        import torch
        self._mask_i8 = self.vec._mask_u8.view(torch.int8)
        if _gym is not None:
            single_obs = _gym.spaces.Dict({
                "action_mask": _gym.spaces.Box(0, 1, shape=(J + 1,), dtype=np.int8),
                "real_obs": _gym.spaces.Box(0.0, 1.0, shape=(J, 7), dtype=np.float32)})
            self.single_observation_space = single_obs
            self.single_action_space = _gym.spaces.Discrete(J + 1)
            self.observation_space = _gym.vector.utils.batch_space(single_obs, num_envs)
            self.action_space = _gym.vector.utils.batch_space(self.single_action_space, num_envs)

    def _out(self):
        obs = {"real_obs": self.vec.real_obs, "action_mask": self._mask_i8}
        if not self.to_numpy:
            return obs
        return {k: v.cpu().numpy() for k, v in obs.items()}

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.vec.seed = int(seed)
        self.vec.reset()
        return self._out(), {}

    def step(self, actions):
        _, reward, done, truncated, info = self.vec.step(actions)
        if self.to_numpy:
            return self._out(), reward.cpu().numpy(), done.cpu().numpy(), truncated.cpu().numpy(), info
        return self._out(), reward, done, truncated, info

    def close(self, **kwargs):
        self.vec.close()


def create_env(config: Dict[str, Any], *args, **kwargs):
    """Mirror of ``JSSEnv/utils.py:32-60``: ``config["env"]`` names the environment ("jss-v1"), the remaining
    keys are its ``env_config``.  ``config["num_envs"] > 1`` returns the batched ``JssGymVectorEnv`` instead
    of the single-env facade (the reference has no batched form)."""
    if not isinstance(config, dict) or "env" not in config:
        raise KeyError("config must be a dict with an 'env' key (JSSEnv/utils.py:41-46)")
    name = config["env"]
    if name not in ("jss-v1", "JssEnv"):
        raise NotImplementedError(f"Environment {name} not recognized (JSSEnv/utils.py:55-58)")
    env_config = config.get("env_config") or {k: v for k, v in config.items() if k not in ("env", "num_envs", "device")}
    if not env_config:
        env_config = None
    n = int(config.get("num_envs", 1))
    if n > 1:
        return JssGymVectorEnv(n, env_config, device=int(config.get("device", 0)), **kwargs)
    from .env import JssEnv
    return JssEnv(env_config, device=int(config.get("device", 0)))
