"""gym.spaces subset: Discrete, MultiDiscrete (with .nvec / .sample()), Box, Dict (GraspingEnv.py:158-167)."""
import numpy as np


class Space:
    def __init__(self, shape=None, dtype=None):
        self.shape, self.dtype = shape, dtype
        self.np_random = np.random.RandomState()

    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]


class Discrete(Space):
    def __init__(self, n):
        super().__init__((), np.int64)
        self.n = int(n)

    def sample(self):
        return int(self.np_random.randint(self.n))

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __repr__(self):
        return f"Discrete({self.n})"


class MultiDiscrete(Space):
    def __init__(self, nvec):
        self.nvec = np.asarray(nvec, dtype=np.int64)
        super().__init__(self.nvec.shape, np.int64)

    def sample(self):
        return (self.np_random.random_sample(self.nvec.shape) * self.nvec).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.nvec.shape and (0 <= x).all() and (x < self.nvec).all()

    def __repr__(self):
        return f"MultiDiscrete({self.nvec})"


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        shape = shape if shape is not None else np.shape(low)
        super().__init__(tuple(shape), dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self.shape)
        self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self.shape)

    def sample(self):
        return self.np_random.uniform(self.low, np.where(np.isfinite(self.high), self.high, 1.0)).astype(self.dtype)

    def __repr__(self):
        return f"Box{self.shape}"


class Dict(Space):
    def __init__(self, spaces):
        super().__init__(None, None)
        self.spaces = dict(spaces)

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}

    def __repr__(self):
        return "Dict(" + ", ".join(f"{k}:{v}" for k, v in self.spaces.items()) + ")"
