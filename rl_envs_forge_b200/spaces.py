"""Minimal gymnasium.spaces look-alikes (gymnasium is not installed in the target image).

The reference exposes `action_space` / `observation_space` built from gymnasium 0.29.1 `Discrete`, `Tuple`, `Box`
(grid_world.py:90-93, k_armed_bandit.py:149-150, cart_pole.py:50-53, pendulum_disk.py:49-55).  When gymnasium is
importable the real classes are used; otherwise these stand-ins provide `n`, `shape`, `dtype`, `low`, `high`,
`contains`, `sample` with the same semantics.
"""
import numpy as np

try:  # pragma: no cover - exercised only where gymnasium exists
    from gymnasium.spaces import Box, Discrete, Tuple  # type: ignore  # noqa: F401
    HAVE_GYMNASIUM = True
except Exception:  # ImportError and friends
    HAVE_GYMNASIUM = False

    class _Space:
        def __init__(self, shape=None, dtype=None, seed=None):
            self._shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)
            self._rng = np.random.default_rng(seed)

        @property
        def shape(self):
            return self._shape

        def __contains__(self, x):
            return self.contains(x)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

    class Discrete(_Space):
        def __init__(self, n, seed=None, start=0):
            super().__init__((), np.int64, seed)
            self.n, self.start = int(n), int(start)

        def contains(self, x):
            if isinstance(x, (int, np.integer)):
                v = int(x)
            elif isinstance(x, np.ndarray) and x.shape == () and np.issubdtype(x.dtype, np.integer):
                v = int(x)
            else:
                return False
            return self.start <= v < self.start + self.n

        def sample(self):
            return int(self.start + self._rng.integers(self.n))

        def __repr__(self):
            return f"Discrete({self.n})"

        def __eq__(self, other):
            return isinstance(other, Discrete) and (self.n, self.start) == (other.n, other.start)

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            dtype = np.dtype(dtype)
            if shape is None:
                shape = np.shape(low) if np.shape(low) != () else np.shape(high)
            super().__init__(shape, dtype, seed)
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), self._shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), self._shape).copy()

        def contains(self, x):
            if not isinstance(x, np.ndarray):
                try:
                    x = np.asarray(x, dtype=self.dtype)
                except (ValueError, TypeError):
                    return False
            return bool(np.can_cast(x.dtype, self.dtype) and x.shape == self.shape
                        and np.all(x >= self.low) and np.all(x <= self.high))

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1e6)
            hi = np.where(np.isfinite(self.high), self.high, 1e6)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

        def __eq__(self, other):
            return (isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype
                    and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high))

    class Tuple(_Space):
        def __init__(self, spaces, seed=None):
            super().__init__(None, None, seed)
            self.spaces = tuple(spaces)

        def contains(self, x):
            if isinstance(x, (list, np.ndarray)):
                x = tuple(x)
            return isinstance(x, tuple) and len(x) == len(self.spaces) and all(
                s.contains(p) for s, p in zip(self.spaces, x))

        def sample(self):
            return tuple(s.sample() for s in self.spaces)

        def __len__(self):
            return len(self.spaces)

        def __getitem__(self, i):
            return self.spaces[i]

        def __repr__(self):
            return "Tuple(" + ", ".join(map(repr, self.spaces)) + ")"

        def __eq__(self, other):
            return isinstance(other, Tuple) and self.spaces == other.spaces


class Batched:
    """Batched view of a single-env space (what gymnasium.vector calls `action_space` of a vector env)."""

    def __init__(self, single, num_envs):
        self.single, self.num_envs = single, int(num_envs)

    @property
    def shape(self):
        s = getattr(self.single, "shape", None)
        if s is None:
            return (self.num_envs, len(self.single))
        return (self.num_envs,) + tuple(s)

    @property
    def dtype(self):
        return getattr(self.single, "dtype", None)

    def sample(self):
        return np.stack([np.asarray(self.single.sample()) for _ in range(self.num_envs)])

    def __repr__(self):
        return f"Batched({self.single!r}, num_envs={self.num_envs})"
