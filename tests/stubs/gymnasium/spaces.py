import numpy as np


class Space:
    shape = None
    dtype = None


class Discrete(Space):
    def __init__(self, n, start=0):
        self.n, self.start, self.shape, self.dtype = int(n), int(start), (), np.dtype(np.int64)

    def contains(self, x):
        return self.start <= int(x) < self.start + self.n


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        self.shape, self.dtype = self.nvec.shape, np.dtype(np.int64)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(((0 <= x) & (x < self.nvec)).all())


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(((x >= self.low) & (x <= self.high)).all())


class Dict(Space):
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def contains(self, x):
        return set(x.keys()) == set(self.spaces.keys()) and all(self.spaces[k].contains(x[k]) for k in x)
