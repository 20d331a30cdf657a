"""Discrete / Box / Dict with the members gobblet.py:140-153 and the wrappers use."""
import numpy as np


class Space:
    def __contains__(self, x):
        return self.contains(x)


class Discrete(Space):
    def __init__(self, n, start=0):
        self.n, self.start = int(n), int(start)
        self.shape, self.dtype = (), np.dtype(np.int64)

    def contains(self, x):
        if isinstance(x, (bool, np.bool_)):
            return False
        if isinstance(x, int):
            v = x
        elif isinstance(x, (np.generic, np.ndarray)) and np.issubdtype(x.dtype, np.integer) and x.shape == ():
            v = int(x)
        else:
            return False
        return self.start <= v < self.start + self.n

    def sample(self, mask=None):
        if mask is not None:
            return int(np.random.choice(np.flatnonzero(mask))) + self.start
        return int(np.random.randint(self.n)) + self.start

    def __eq__(self, other):
        return isinstance(other, Discrete) and (self.n, self.start) == (other.n, other.start)

    def __repr__(self):
        return f"Discrete({self.n})"


class Box(Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return bool(
            np.can_cast(x.dtype, self.dtype) and x.shape == self.shape
            and np.all(x >= self.low) and np.all(x <= self.high)
        )

    def sample(self):
        return np.random.randint(self.low, self.high + 1).astype(self.dtype)

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


class Dict(Space):
    def __init__(self, spaces):
        self.spaces = dict(spaces)

    def __getitem__(self, k):
        return self.spaces[k]

    def keys(self):
        return self.spaces.keys()

    def contains(self, x):
        return isinstance(x, dict) and x.keys() == self.spaces.keys() and all(
            self.spaces[k].contains(x[k]) for k in self.spaces
        )

    def sample(self):
        return {k: s.sample() for k, s in self.spaces.items()}

    def __repr__(self):
        return "Dict(" + ", ".join(f"{k!r}: {s}" for k, s in self.spaces.items()) + ")"
