"""Shared helpers of the parity tests: seeded action streams (the reference's two input distributions)."""
import numpy as np

SIZES = [3, 3, 3, 2, 2, 3]


def encode(heads):
    """MegaverseGym::setActions (megaverse.cpp:100-116)"""
    mask, idx = 0, 0
    for a, sz in zip(heads, SIZES):
        if a > 0:
            mask |= 1 << (idx + int(a))
        idx += sz - 1
    return mask


def random_bit_actions(rng, n):
    """megaverse_test_app.cpp:140-147: one uniformly random action bit per agent (bit 0 = no-op)"""
    return (1 << rng.integers(0, 11, size=n)).astype(np.int32)


def random_head_actions(rng, n):
    """action_space.sample(): independent uniform sample of each Discrete head (megaverse/tests/test_env.py:12-13)"""
    out = np.zeros(n, dtype=np.int32)
    for i in range(n):
        out[i] = encode([rng.integers(0, s) for s in SIZES])
    return out


def purposeful_actions(rng, n, step):
    """mostly walk forward / turn, interact often: makes pick-ups, placements and rewards actually happen"""
    out = np.zeros(n, dtype=np.int32)
    for i in range(n):
        r = rng.random()
        m = 0
        if r < 0.7:
            m |= 1 << 3  # forward
        if rng.random() < 0.25:
            m |= (1 << 5) if rng.random() < 0.5 else (1 << 6)
        if rng.random() < 0.3:
            m |= 1 << 8  # interact
        if rng.random() < 0.1:
            m |= 1 << 7  # jump
        if rng.random() < 0.1:
            m |= (1 << 9) if rng.random() < 0.6 else (1 << 10)
        out[i] = m
    return out
