import numpy as np

from .. import spaces


def batch_space(space, n):
    if isinstance(space, spaces.Box):
        out = spaces.Box(0, 0, shape=(n,) + space.shape, dtype=space.dtype)
        out.low = np.broadcast_to(space.low, out.shape).copy()
        out.high = np.broadcast_to(space.high, out.shape).copy()
        return out
    if isinstance(space, spaces.Discrete):
        return spaces.MultiDiscrete(np.full((n,), space.n))
    if isinstance(space, spaces.Dict):
        return spaces.Dict({k: batch_space(v, n) for k, v in space.spaces.items()})
    raise TypeError(type(space))
