#!/usr/bin/env python
"""Host submission rate of mv_step_device vs the GPU's execution rate (is the async loop CPU-bound?)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from megaverse_b200 import capi

E = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = capi.Engine("TowerBuilding", E, 1, 128, 72, num_threads=8)
for e in range(E):
    g.seed_env(e, 42 + e)
g.reset()
K = 3000
acts = torch.from_numpy((1 << np.random.default_rng(1).integers(0, 11, size=(K, E))).astype(np.int32)).cuda()
torch.cuda.synchronize()
for ov in (1, 0):
    g.set_option("overlap", ov)
    for t in range(100):
        g.step_device(acts.data_ptr() + t * E * 4)
    g.sync()
    t0 = time.perf_counter()
    for t in range(K):
        g.step_device(acts.data_ptr() + t * E * 4)
    t1 = time.perf_counter()
    g.sync()
    t2 = time.perf_counter()
    print("overlap=%d: submit loop %.1f us/step, until drained %.1f us/step" % (ov, (t1 - t0) / K * 1e6, (t2 - t0) / K * 1e6))
g.close()
