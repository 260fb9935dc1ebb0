#!/usr/bin/env python
"""device-resident rate with and without option "cull" (instance-level frustum culling in the geometry kernel)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from megaverse_b200 import capi

K = 600
for scenario, E, A in (("TowerBuilding", 256, 1), ("HexExplore", 256, 2), ("Collect", 256, 4), ("ObstaclesHard", 256, 1)):
    for cull in (0, 1, 0, 1):
        g = capi.Engine(scenario, E, A, 128, 72, num_threads=8)
        g.set_option("cull", cull)
        for e in range(E):
            g.seed_env(e, 42 + e)
        g.reset()
        acts = torch.from_numpy((1 << np.random.default_rng(1).integers(0, 11, size=(K + 100, E * A))).astype(np.int32)).cuda()
        torch.cuda.synchronize()
        for t in range(100):
            g.step_device(acts.data_ptr() + t * E * A * 4)
        g.sync()
        t0 = time.perf_counter()
        for t in range(100, 100 + K):
            g.step_device(acts.data_ptr() + t * E * A * 4)
        g.sync()
        dt = time.perf_counter() - t0
        s_ms, r_ms = g.last_kernel_ms()
        print(f"{scenario:14s} {E}x{A} cull={cull}: {E * A * K / dt / 1e6:.3f} M obs/s  ({dt / K * 1e6:.1f} us/step)", flush=True)
        g.close()
