"""TEST-ONLY minimal stand-in for the `gymnasium` package (absent from the build image and the GPU box).

Only what the drop-in boundary touches (reference: JSSEnv/__init__.py:6-9 `register`, README.md:46 `gym.make`,
jss_env.py:8,14,97,112-119 `gym.Env` / `gym.spaces`): a registry with `register` / `make`, `Env`, the three
space classes with shape/dtype/contains, and `vector.VectorEnv` + `vector.utils.batch_space`.  It lives under
tests/stubs and is put on sys.path ONLY by the subprocess-based tests in tests/test_gym_boundary.py."""
import importlib

from . import spaces  # noqa: F401
from .envs.registration import register, registry  # noqa: F401

__version__ = "0.0-stub"


class Env:
    metadata = {}
    action_space = None
    observation_space = None

    def reset(self, *, seed=None, options=None):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def close(self):
        pass

    @property
    def unwrapped(self):
        return self


def make(id, **kwargs):
    """gymnasium.make: resolve `id` in the registry, import "module:Class", instantiate with kwargs."""
    if id not in registry:
        raise KeyError(f"No registered env with id: {id}")
    spec = registry[id]
    entry = spec["entry_point"]
    if isinstance(entry, str):
        mod, _, attr = entry.partition(":")
        entry = getattr(importlib.import_module(mod), attr)
    return entry(**{**spec.get("kwargs", {}), **kwargs})


from . import vector  # noqa: E402,F401
