#!/usr/bin/env python
"""Throughput and kernel times of other BASELINE configs (device-resident loop): scenario, envs, agents, depth."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from megaverse_b200 import capi

CONFIGS = [("TowerBuilding", 256, 1, False), ("TowerBuilding", 4096, 1, False), ("ObstaclesHard", 2048, 1, True), ("Collect", 1024, 4, False)]
if len(sys.argv) > 1:
    CONFIGS = [(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), len(sys.argv) > 4 and sys.argv[4] == "depth")]
for scenario, E, A, depth in CONFIGS:
    g = capi.Engine(scenario, E, A, 128, 72, num_threads=16, depth=depth)
    for k, v in [kv.split("=") for kv in os.environ.get("MV_OPTS", "").split(",") if kv]:
        g.set_option(k, int(v))
    for e in range(E):
        g.seed_env(e, 42 + e)
    g.reset()
    K = 300
    acts = torch.from_numpy((1 << np.random.default_rng(1).integers(0, 11, size=(K, E * A))).astype(np.int32)).cuda()
    torch.cuda.synchronize()
    for t in range(50):
        g.step_device(acts.data_ptr() + t * E * A * 4)
    g.sync()
    t0 = time.perf_counter()
    for t in range(K):
        g.step_device(acts.data_ptr() + t * E * A * 4)
    g.sync()
    dt = (time.perf_counter() - t0) / K
    g.set_option("overlap", 0)
    ks = []
    for t in range(20):
        g.step_device(acts.data_ptr() + t * E * A * 4)
        g.sync()
        ks.append(g.last_kernel_ms())
    ks = np.array(ks[5:]).mean(axis=0)
    g.raster_stats(True, False)
    for t in range(10):
        g.step_device(acts.data_ptr() + t * E * A * 4)
    g.sync()
    st = g.raster_stats(False, True)
    print("   per work item:", {k: round(v / max(1, st["work_items"]), 1) for k, v in st.items() if k != "work_items"})
    print("%-14s E=%-5d A=%d depth=%d: %.3f ms/step = %.2fM obs/s; step kernel %.3f ms, raster %.3f ms; faults %d" % (scenario, E, A, depth, dt * 1e3, E * A / dt / 1e6, ks[0], ks[1], g.faults()), g.raster_config())
    g.close()
