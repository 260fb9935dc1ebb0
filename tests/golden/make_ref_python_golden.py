"""Generates tests/golden/ref_python_env_golden.npz by driving the REFERENCE's own Python class (megaverse/megaverse_env.py, imported from
/root/reference) on the reference's own pybind module and env library (compiled in place on the Bullet stand-in with null renderers,
see tests/refpy.py and oracle/ref_shim/binding_shim.cpp): MegaverseEnv(...).seed(s); reset(); step(actions) with six-head actions, one
reward-shaping change on the way.  Stored: the actions and everything step() returned except the (blank) frames -- rewards, dones, the
true_reward infos -- and the reward shaping dictionaries.  Runs in the build container only; the fixture travels.

    python tests/golden/make_ref_python_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault("BOXOBAN_LEVELS", os.path.join(HERE, "boxoban"))
import helpers  # noqa: E402
import refpy  # noqa: E402

CASES = [  # name, scenario, envs, agents, seed, ticks, params, (tick, actor, shaping-key, value) or None
    ("towerbuilding", "TowerBuilding", 4, 2, 11, 240, None, (60, 3, "towerPickedUpObject", 0.7)),
    ("collect", "Collect", 3, 2, 12, 240, {"episodeLengthSec": -2.0}, (40, 1, "collectSingleGood", 2.5)),
    ("obstacleseasy", "ObstaclesEasy", 4, 1, 13, 240, None, None),
    ("hexmemory", "HexMemory", 2, 2, 14, 160, None, None),
    ("sokoban", "Sokoban", 3, 1, 15, 200, {"episodeLengthSec": 8.0}, None),
]


def main():
    RefEnv = refpy.reference_env_class()
    out = {}
    for name, scenario, E, A, seed, T, params, change in CASES:
        env = RefEnv(scenario, num_envs=E, num_agents_per_env=A, num_simulation_threads=1, use_vulkan=True, params=params)
        N = E * A
        env.seed(seed)
        env.reset()
        rng = np.random.default_rng(seed)
        acts = np.zeros((T, N, 6), np.int32)
        rew = np.zeros((T, N), np.float32)
        done = np.zeros((T, N), np.uint8)
        true_reward = np.full((T, N), np.nan, np.float32)
        for t in range(T):
            if change and t == change[0]:
                rs = dict(env.get_current_reward_shaping(change[1]))
                rs[change[2]] = change[3]
                env.set_reward_shaping(rs, change[1])
            # purposeful masks expressed as six-head tuples (the reference's action format)
            for i in range(N):
                m = int(helpers.purposeful_actions(rng, 1, t)[0])
                heads = [0] * 6
                heads[0] = 1 if m & 2 else (2 if m & 4 else 0)
                heads[1] = 1 if m & 8 else (2 if m & 16 else 0)
                heads[2] = 1 if m & 32 else (2 if m & 64 else 0)
                heads[3] = 1 if m & 128 else 0
                heads[4] = 1 if m & 256 else 0
                heads[5] = 1 if m & 512 else (2 if m & 1024 else 0)
                acts[t, i] = heads
            _, r, d, infos = env.step([list(map(int, a)) for a in acts[t]])
            rew[t], done[t] = np.asarray(r, np.float32), np.asarray(d, np.uint8)
            for i, inf in enumerate(infos):
                if inf:
                    true_reward[t, i] = inf["true_reward"]
        shaping0 = env.get_default_reward_shaping()
        final = [env.get_current_reward_shaping(i) for i in range(N)]
        env.close()
        out[name + "/meta"] = np.array([E, A, seed, T], np.int32)
        out[name + "/scenario"] = np.array(scenario)
        out[name + "/param_keys"] = np.array(list((params or {}).keys()), dtype="U32")
        out[name + "/param_vals"] = np.array(list((params or {}).values()), np.float64)
        out[name + "/change"] = np.array([change[0], change[1]] if change else [-1, -1], np.int32)
        out[name + "/change_key"] = np.array(change[2] if change else "")
        out[name + "/change_val"] = np.array(change[3] if change else 0.0, np.float64)
        out[name + "/actions"], out[name + "/rewards"], out[name + "/dones"], out[name + "/true_reward"] = acts, rew, done, true_reward
        out[name + "/shaping_keys"] = np.array(sorted(shaping0), dtype="U64")
        out[name + "/shaping_default"] = np.array([shaping0[k] for k in sorted(shaping0)], np.float32)
        out[name + "/shaping_final"] = np.array([[f[k] for k in sorted(shaping0)] for f in final], np.float32)
        print(name, "episodes finished:", int(done[:, ::A].sum()), "reward events:", int((rew != 0).sum()))
    path = os.path.join(HERE, "ref_python_env_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
