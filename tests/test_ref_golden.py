"""Golden vectors produced by the REFERENCE's own env library (tests/golden/make_ref_golden.py: /root/reference's env.cpp, agent.cpp,
character controller and scenario sources on the Bullet stand-in, see oracle/ref_shim/env_shim.cpp) replayed on the oracle (CPU
suite) and on the CUDA engine through the C ABI (GPU suite).  Everything is compared bit for bit: rewards, done flags, true
objectives at episode ends, and after every tick the number of drawables and a CRC-32 over the complete drawable list
([mesh type, 24-bit colour, the 16 floats of the absolute matrix] per drawable, in draw order) of every env."""
import os
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "ref_env_golden.npz")
# env/const.hpp's palette in the product's / oracle's index order (pinned against the reference by tests/test_ref_shim.py)
PALETTE = [0xffdd3c, 0x3bb372, 0x50c878, 0x2eb5d0, 0xadd8e6, 0x3a7fa6, 0x2c3e50, 0xffb400, 0xb3b3b3, 0x555555, 0x222222,
           0xffffff, 0xff0000, 0xffa770, 0xd468ee, 0xffe6e6, 0xffffe6, 0xccffcc, 0xe6ecff, 0xd9d9d9, 0xf2e6ff, 0xffebcc]
CASES = ["TowerBuilding", "ObstaclesHard", "Collect", "Sokoban", "Rearrange", "HexExplore", "HexMemory", "TowerBuilding_8agents_widepitch",
         "ObstaclesEasy_custom_course", "ObstaclesLava", "Collect_8agents_short"]


def _crc(inst18):
    """drawable list as the reference shim lays it out: [mesh, 0xRRGGBB, 16 matrix bit patterns] uint32 rows"""
    inst18 = np.ascontiguousarray(inst18, np.float32).reshape(-1, 18)
    rows = np.zeros((len(inst18), 18), np.uint32)
    rows[:, 0] = inst18[:, 0].astype(np.uint32)
    rows[:, 1] = np.asarray(PALETTE, np.uint32)[inst18[:, 1].astype(np.int64)]
    rows[:, 2:] = inst18[:, 2:].copy().view(np.uint32)
    return len(rows), zlib.crc32(rows.tobytes())


def _replay(make, case):
    g = np.load(GOLDEN)
    scenario = str(g[case + "/scenario"])
    A, E, T = (int(v) for v in g[case + "/meta"])
    params = {str(k): float(v) for k, v in zip(g[case + "/param_keys"], g[case + "/param_vals"])}
    sim = make(scenario, E, A, params)
    for e in range(E):
        sim.seed_env(e, 1000 + e)
    sim.reset()
    acts, rew, tobj, done = g[case + "/actions"], g[case + "/rewards"], g[case + "/true_objectives"], g[case + "/dones"]
    ninst, crc = g[case + "/n_inst"], g[case + "/crc"]

    def check_drawables(t):
        for e in range(E):
            n, c = _crc(sim.instances(e))
            assert n == ninst[t, e], "%s: drawable count of env %d after tick %d: %d, reference %d" % (scenario, e, t - 1, n, ninst[t, e])
            assert c == crc[t, e], "%s: drawable list of env %d after tick %d differs from the reference's" % (scenario, e, t - 1)

    check_drawables(0)
    for t in range(T):
        sim.step(acts[t].reshape(-1))
        r = np.array(sim.rewards(), np.float32).reshape(E, A)
        assert np.array_equal(r.view(np.uint32), rew[t].view(np.uint32)), "%s: rewards at tick %d: %s, reference %s" % (scenario, t, r, rew[t])
        d = np.array(sim.dones(), np.uint8).reshape(E)
        assert np.array_equal(d, done[t]), "%s: dones at tick %d: %s, reference %s" % (scenario, t, d, done[t])
        if d.any():
            to = np.array(sim.true_objectives(), np.float32).reshape(E, A)
            assert np.array_equal(to[d != 0], tobj[t][d != 0]), "%s: true objectives at tick %d" % (scenario, t)
        check_drawables(t + 1)
    sim.close()
    return int(done.sum()), int((rew != 0).sum())


@pytest.mark.parametrize("case", CASES)
def test_oracle_replays_the_reference_env_library_golden(built, case):
    import orc

    _replay(lambda s, E, A, p: orc.Oracle(s, E, A, params=p, render=False), case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_device_replays_the_reference_env_library_golden(built, case):
    from megaverse_b200 import capi

    def make(s, E, A, p):
        g = capi.Engine(s, E, A, 128, 72, num_threads=2, params=p)
        return g

    _replay(make, case)
    # the same through the asynchronous device-resident path is covered against the oracle in test_parity_gpu.py
