"""Observation / action spaces.  gymnasium's own classes are used when it is installed
(gobblet.py:140-153 builds Discrete(54) and a Dict of two int8 Boxes); otherwise these minimal
equivalents provide the members PettingZoo / Tianshou callers touch (contains, sample, shape, dtype, n)."""
import numpy as np

try:  # pragma: no cover - gymnasium is not in the build image
    from gymnasium.spaces import Box, Dict, Discrete  # noqa: F401
except ImportError:

    class _Space:
        def __contains__(self, x):
            return self.contains(x)

    class Discrete(_Space):
        def __init__(self, n):
            self.n = int(n)
            self.shape = ()
            self.dtype = np.dtype(np.int64)

        def contains(self, x):
            if isinstance(x, (bool, np.bool_)):
                return False
            if isinstance(x, (np.generic, np.ndarray)):
                if not (np.issubdtype(x.dtype, np.integer) and x.shape == ()):
                    return False
            elif not isinstance(x, int):
                return False
            return 0 <= int(x) < self.n

        def sample(self, mask=None):
            if mask is not None:
                return int(np.random.choice(np.flatnonzero(mask)))
            return int(np.random.randint(self.n))

        def __eq__(self, other):
            return isinstance(other, Discrete) and other.n == self.n

        def __repr__(self):
            return f"Discrete({self.n})"

    class Box(_Space):
        def __init__(self, low, high, shape, dtype):
            self.shape = tuple(shape)
            self.dtype = np.dtype(dtype)
            self.low = np.full(self.shape, low, self.dtype)
            self.high = np.full(self.shape, high, self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return bool(x.shape == self.shape and np.can_cast(x.dtype, self.dtype)
                        and (x >= self.low).all() and (x <= self.high).all())

        def sample(self):
            return np.random.randint(self.low, self.high + 1).astype(self.dtype)

        def __repr__(self):
            return f"Box({self.low.flat[0]}, {self.high.flat[0]}, {self.shape}, {self.dtype})"

    class Dict(_Space):
        def __init__(self, spaces):
            self.spaces = dict(spaces)

        def __getitem__(self, key):
            return self.spaces[key]

        def keys(self):
            return self.spaces.keys()

        def contains(self, x):
            return (isinstance(x, dict) and set(x) == set(self.spaces)
                    and all(s.contains(x[k]) for k, s in self.spaces.items()))

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k!r}: {v!r}" for k, v in self.spaces.items()) + ")"
