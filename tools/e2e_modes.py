#!/usr/bin/env python
"""End-to-end step time through host buffers (mv_set_actions + mv_step) under the delivery modes: zero-copy stores from the
raster kernel vs one copy-engine download after it."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megaverse_b200 import capi

E = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = capi.Engine("TowerBuilding", E, 1, 128, 72, num_threads=8)
for e in range(E):
    g.seed_env(e, 42 + e)
g.reset()
rng = np.random.default_rng(1)
acts = (1 << rng.integers(0, 11, size=(600, E))).astype(np.int32)
for name, opts in [("zero_copy", {"zero_copy": 1}), ("one copy after the raster", {"zero_copy": 0})]:
    for k, v in opts.items():
        g.set_option(k, v)
    for t in range(50):
        g.step(acts[t])
    t0 = time.perf_counter()
    for t in range(50, 550):
        g.step(acts[t])
    dt = (time.perf_counter() - t0) / 500
    print("%-26s %.1f us/step = %.2fM obs/s (faults %d)" % (name, dt * 1e6, E / dt / 1e6, g.faults()))
g.close()
