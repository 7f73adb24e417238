"""Spaces: gymnasium's when it is importable, otherwise minimal look-alikes (gymnasium is absent from this image).

Only what the reference's envs expose is mirrored (envs/robot_env.py:87-100): Box and Dict, plus batching.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gymnasium is not installed in the build container
    from gymnasium.spaces import Box, Dict  # type: ignore
    from gymnasium.vector.utils import batch_space  # type: ignore
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.dtype = np.dtype(dtype)
            self.shape = tuple(shape) if shape is not None else np.shape(low)
            self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1e3)
            hi = np.where(np.isfinite(self.high), self.high, 1e3)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    class Dict:
        def __init__(self, spaces):
            self.spaces = dict(spaces)

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x):
            return isinstance(x, dict) and x.keys() == self.spaces.keys() and all(s.contains(x[k]) for k, s in self.spaces.items())

        def __repr__(self):
            return f"Dict({self.spaces})"

    def batch_space(space, n):
        if isinstance(space, Dict):
            return Dict({k: batch_space(s, n) for k, s in space.spaces.items()})
        return Box(np.broadcast_to(space.low, (n,) + space.shape), np.broadcast_to(space.high, (n,) + space.shape),
                   shape=(n,) + space.shape, dtype=space.dtype)
