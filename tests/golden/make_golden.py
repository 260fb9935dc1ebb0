"""Generates tests/golden/tower_golden.npz from the ORACLE (oracle/liborc.so).

The reference itself cannot run in the build container (Bullet 2.89, Vulkan and EGL are absent: SURVEY.md 8c) and its own
tests hold no trajectory / reward / pixel golden vectors, so these goldens pin the ORACLE's behaviour over time (any change
of the restatement's arithmetic shows up as a golden diff); the pieces of the reference that DO build here (Magnum math and
primitives) are pinned separately by tests/test_ref_shim.py.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402
import orc  # noqa: E402


def main():
    E, A, seed, steps = 3, 2, 2024, 240
    o = orc.Oracle("TowerBuilding", E, A)
    o.seed(seed)
    o.reset()
    first = o.obs()[0].copy()
    rng = np.random.default_rng(5)
    actions, rewards, dones = [], [], []
    for t in range(steps):
        a = helpers.purposeful_actions(rng, E * A, t)
        o.step(a)
        actions.append(a)
        rewards.append(o.rewards())
        dones.append(o.dones())
    np.savez_compressed(os.path.join(HERE, "tower_golden.npz"), E=E, A=A, seed=seed, actions=np.array(actions), rewards=np.array(rewards),
                        dones=np.array(dones), first_frame=first, last_frame=o.obs()[0].copy(), final_state0=o.state(0))
    print("rewards earned:", float(np.abs(np.array(rewards)).sum()))


if __name__ == "__main__":
    main()
