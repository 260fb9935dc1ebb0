"""Generates tests/golden/tower_golden.npz from the ORACLE (oracle/liborc.so).

These goldens pin the ORACLE's behaviour over time, frames included (any change of the restatement's arithmetic shows up as a golden
diff).  The reference's own outputs are pinned elsewhere: tests/test_ref_shim.py runs the oracle beside the reference's env library
compiled in place, and make_ref_golden.py / make_ref_python_golden.py write goldens FROM that build (no frames: the reference's
renderers need Vulkan / OpenGL, which this image does not have).

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


SCENARIOS = ["ObstaclesHard", "ObstaclesLava", "Collect", "Sokoban", "Rearrange", "HexExplore", "HexMemory"]


def scenario_run(name, E=3, A=2, seed=77, steps=200):
    """level dumps at reset, then a fixed action stream: rewards, dones and the final states"""
    o = orc.Oracle(name, E, A, render=False)
    o.seed(seed)
    o.reset()
    levels = [o.level(e) for e in range(E)]
    rng = np.random.default_rng(11)
    rewards, dones = [], []
    for t in range(steps):
        o.step(helpers.purposeful_actions(rng, E * A, t))
        rewards.append(o.rewards().copy()); dones.append(o.dones().copy())
    states = [o.state(e) for e in range(E)]
    o.close()
    return levels, np.array(rewards), np.array(dones), states


def main_scenarios():
    os.environ.setdefault("BOXOBAN_LEVELS", os.path.join(HERE, "boxoban"))
    out = {}
    for name in SCENARIOS:
        levels, rewards, dones, states = scenario_run(name)
        for e, lv in enumerate(levels):
            out["%s_level%d" % (name, e)] = lv
        out[name + "_rewards"] = rewards
        out[name + "_dones"] = dones
        for e, st in enumerate(states):
            out["%s_state%d" % (name, e)] = st
    np.savez_compressed(os.path.join(HERE, "scenarios_golden.npz"), **out)
    print("scenarios:", {n: float(np.abs(out[n + "_rewards"]).sum()) for n in SCENARIOS})


if __name__ == "__main__":
    main()
    main_scenarios()
