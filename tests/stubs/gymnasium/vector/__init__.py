import enum

from . import utils  # noqa: F401


class AutoresetMode(enum.Enum):
    NEXT_STEP = "NextStep"
    SAME_STEP = "SameStep"
    DISABLED = "Disabled"


class VectorEnv:
    metadata = {}
    num_envs = None
    single_observation_space = None
    single_action_space = None
    observation_space = None
    action_space = None

    def reset(self, *, seed=None, options=None):
        raise NotImplementedError

    def step(self, actions):
        raise NotImplementedError

    def close(self, **kwargs):
        pass
