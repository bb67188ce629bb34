"""gym.spaces if gym (or gymnasium) is installed, else a minimal stand-in with the same attributes.
The reference builds spaces.Tuple of spaces.Box / spaces.Discrete (e.g. PredatorCapturePrey.py:47-56)."""
import numpy as np

try:                                            # pragma: no cover - depends on the environment
    from gym.spaces import Box, Discrete, Tuple
except Exception:
    try:                                        # pragma: no cover
        from gymnasium.spaces import Box, Discrete, Tuple
    except Exception:
        class Discrete(object):
            def __init__(self, n):
                self.n = int(n)
                self.shape = ()
                self.dtype = np.int64

            def sample(self):
                return int(np.random.randint(self.n))

            def contains(self, x):
                return 0 <= int(x) < self.n

            def __repr__(self):
                return "Discrete(%d)" % self.n

        class Box(object):
            def __init__(self, low, high, shape=None, dtype=np.float32):
                self.shape = tuple(shape)
                self.dtype = np.dtype(dtype)
                self.low = np.full(self.shape, low, dtype=self.dtype)
                self.high = np.full(self.shape, high, dtype=self.dtype)

            def sample(self):
                return np.random.uniform(self.low, self.high).astype(self.dtype)

            def contains(self, x):
                x = np.asarray(x)
                return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

            def __repr__(self):
                return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)

        class Tuple(object):
            def __init__(self, spaces):
                self.spaces = tuple(spaces)

            def sample(self):
                return tuple(s.sample() for s in self.spaces)

            def __len__(self):
                return len(self.spaces)

            def __getitem__(self, i):
                return self.spaces[i]

            def __iter__(self):
                return iter(self.spaces)

            def __repr__(self):
                return "Tuple(%s)" % ", ".join(map(repr, self.spaces))
