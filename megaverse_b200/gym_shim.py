"""Minimal stand-ins for the `gym` names megaverse_env.py uses (gym.Env, spaces.Discrete / Tuple / Box).
`gym` / `gymnasium` are used when installed; this shim only exists so the env class keeps its surface without them."""
import numpy as np

try:  # pragma: no cover - depends on the installation
    import gym  # type: ignore
    from gym import spaces  # type: ignore

    Env, Discrete, Tuple, Box = gym.Env, spaces.Discrete, spaces.Tuple, spaces.Box
except Exception:  # noqa: BLE001
    try:  # pragma: no cover
        import gymnasium as gym  # type: ignore
        from gymnasium import spaces  # type: ignore

        Env, Discrete, Tuple, Box = gym.Env, spaces.Discrete, spaces.Tuple, spaces.Box
    except Exception:  # noqa: BLE001

        class Env:  # noqa: D401
            """gym.Env placeholder"""

            metadata = {}

        class Discrete:
            def __init__(self, n):
                self.n = int(n)
                self._rng = np.random.default_rng()

            def sample(self):
                return int(self._rng.integers(0, self.n))

            def contains(self, x):
                return 0 <= int(x) < self.n

            def __repr__(self):
                return "Discrete(%d)" % self.n

        class Tuple:
            def __init__(self, spaces):
                self.spaces = tuple(spaces)

            def sample(self):
                return tuple(s.sample() for s in self.spaces)

            def __len__(self):
                return len(self.spaces)

            def __getitem__(self, i):
                return self.spaces[i]

            def __repr__(self):
                return "Tuple(%s)" % ", ".join(map(repr, self.spaces))

        class Box:
            def __init__(self, low, high, shape, dtype=np.float32):
                self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)

            def __repr__(self):
                return "Box(%s, %s, %s, %s)" % (self.low, self.high, self.shape, self.dtype)
