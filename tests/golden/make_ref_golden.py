"""Generates tests/golden/ref_env_golden.npz from the REFERENCE's own env library (oracle/_ref/libmvenv.so: env.cpp, agent.cpp, the
character controller and every scenario source of /root/reference compiled in place on the Bullet stand-in of
oracle/ref_shim/mini_bullet -- see oracle/ref_shim/env_shim.cpp for what is real and what is stand-in).  Needs /root/reference (to
build that library), so it runs in the build container only; the fixture travels.

For every case: E reference envs (Env::seed(1000 + i)) inside the reference's own VectorEnv (vector_env.cpp, compiled as is, null
renderer), read the way MegaverseGym reads it (megaverse.cpp:118-143), T ticks of scripted actions.  Stored per tick and env: the
action masks, rewards, done flags, true objectives, the number of drawables and a CRC-32 of the whole drawable list after the tick
([mesh type, 24-bit colour, 16 matrix bit patterns] per drawable, draw order).  tests replay the actions on the oracle (CPU suite) and
on the device engine (GPU suite) and compare all of it bit for bit.

    python tests/golden/make_ref_golden.py
"""
import ctypes as C
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE))
os.environ.setdefault("BOXOBAN_LEVELS", os.path.join(HERE, "boxoban"))
import helpers  # noqa: E402

MAZE_SEED_XOR = 0x6D617A65
CASES = [  # case name, scenario, agents, envs, ticks, params
    ("TowerBuilding", "TowerBuilding", 2, 4, 240, {"episodeLengthSec": -33.0}),
    ("ObstaclesHard", "ObstaclesHard", 1, 4, 240, {}),
    ("Collect", "Collect", 2, 4, 240, {"episodeLengthSec": -2.0}),
    ("Sokoban", "Sokoban", 1, 3, 200, {"episodeLengthSec": 8.0}),
    ("Rearrange", "Rearrange", 2, 3, 200, {"episodeLengthSec": 8.0}),
    ("HexExplore", "HexExplore", 1, 2, 160, {"episodeLengthSec": 6.0}),
    ("HexMemory", "HexMemory", 2, 2, 160, {"episodeLengthSec": 4.0}),
    ("TowerBuilding_8agents_widepitch", "TowerBuilding", 8, 2, 200, {"verticalLookLimitRad": 0.9}),
    ("ObstaclesEasy_custom_course", "ObstaclesEasy", 2, 3, 240, {"obstaclesMinNumPlatforms": 3, "obstaclesMaxNumPlatforms": 5, "obstaclesMinGap": 2, "obstaclesMaxGap": 4,
                                                                "obstaclesMinLava": 2, "obstaclesMaxLava": 6, "obstaclesMinHeight": 1, "obstaclesMaxHeight": 4,
                                                                "obstaclesNumAllowedMaxDifficulty": 2}),
    ("ObstaclesLava", "ObstaclesLava", 2, 3, 200, {}),
    ("Collect_8agents_short", "Collect", 8, 2, 200, {"episodeLengthSec": -4.0}),
]


def load():
    lib = os.path.join(ROOT, "oracle", "_ref", "libmvenv.so")
    import subprocess

    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "envlib"])
    R = C.CDLL(lib)
    R.ref_vec_create.restype = C.c_void_p
    R.ref_vec_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int, C.c_uint]
    R.ref_vec_destroy.argtypes = [C.c_void_p]
    R.ref_vec_seed_env.argtypes = [C.c_void_p, C.c_int, C.c_int]
    R.ref_vec_reset.argtypes = [C.c_void_p]
    R.ref_vec_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    R.ref_vec_dump.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return R


def drawables(R, v, e, buf):
    n = R.ref_vec_dump(v, e, buf.ctypes.data, len(buf))
    assert n > 0
    d = buf[:n]
    A = int(d[3])
    i = 4 + 3 * A
    n_inst = int(d[i])
    return n_inst, zlib.crc32(d[i + 1:i + 1 + 18 * n_inst].tobytes())


def main():
    R = load()
    buf = np.zeros(1 << 20, np.uint32)
    out = {}
    for key, scenario, A, E, T, params in CASES:
        keys = (C.c_char_p * max(1, len(params)))(*[k.encode() for k in params])
        vals = (C.c_float * max(1, len(params)))(*[float(v) for v in params.values()])
        # the reference's VectorEnv over E envs (vector_env.cpp compiled as is), seeded per env like megaverse_test_app.cpp:250-254
        v = R.ref_vec_create(scenario.encode(), E, A, keys, vals, len(params), MAZE_SEED_XOR)
        for e in range(E):
            R.ref_vec_seed_env(v, e, 1000 + e)
        R.ref_vec_reset(v)
        rng = np.random.default_rng(77)
        acts = np.zeros((T, E, A), np.int32)
        rew = np.zeros((T, E, A), np.float32)
        tobj = np.zeros((T, E, A), np.float32)
        done = np.zeros((T, E), np.uint8)
        ninst = np.zeros((T + 1, E), np.int32)
        crc = np.zeros((T + 1, E), np.uint32)
        for e in range(E):
            ninst[0, e], crc[0, e] = drawables(R, v, e, buf)
        for t in range(T):
            a = np.ascontiguousarray(np.asarray(helpers.purposeful_actions(rng, E * A, t), np.int32).reshape(E, A))
            acts[t] = a
            r, d, o = np.zeros(E * A, np.float32), np.zeros(E, np.uint8), np.zeros(E * A, np.float32)
            # VectorEnv::step, then what MegaverseGym hands out (megaverse.cpp:118-143): last rewards (a finished env was reset inside
            # step(), which zeroes them), done flags, the true objectives VectorEnv kept from the episode's last frame
            R.ref_vec_step(v, a.ctypes.data, r.ctypes.data, d.ctypes.data, o.ctypes.data)
            rew[t], done[t] = r.reshape(E, A), d
            tobj[t][d != 0] = o.reshape(E, A)[d != 0]
            for e in range(E):
                ninst[t + 1, e], crc[t + 1, e] = drawables(R, v, e, buf)
        R.ref_vec_destroy(v)
        out[key + "/scenario"] = np.array(scenario)
        out[key + "/meta"] = np.array([A, E, T], np.int32)
        out[key + "/param_keys"] = np.array(list(params.keys()), dtype="U32")
        out[key + "/param_vals"] = np.array(list(params.values()), np.float32)
        out[key + "/actions"], out[key + "/rewards"], out[key + "/true_objectives"] = acts, rew, tobj
        out[key + "/dones"], out[key + "/n_inst"], out[key + "/crc"] = done, ninst, crc
        print(key, "episodes finished:", int(done.sum()), "reward events:", int((rew != 0).sum()), "drawables:", ninst.min(), "..", ninst.max())
    path = os.path.join(HERE, "ref_env_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
