#!/usr/bin/env python
"""Do several smaller engines on their own streams beat one big one?  (kernels of different sub-batches overlap on the GPU)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from megaverse_b200 import capi

scenario = sys.argv[1] if len(sys.argv) > 1 else "TowerBuilding"
total = int(sys.argv[2]) if len(sys.argv) > 2 else 256
A = int(sys.argv[3]) if len(sys.argv) > 3 else 1
K = 2000
for S in (1, 2, 4, 8):
    E = total // S
    engs = []
    for s in range(S):
        g = capi.Engine(scenario, E, A, 128, 72, num_threads=max(1, 8 // S))
        for e in range(E):
            g.seed_env(e, 42 + s * E + e)
        g.reset()
        engs.append(g)
    acts = torch.from_numpy((1 << np.random.default_rng(1).integers(0, 11, size=(K, total * A))).astype(np.int32)).cuda()
    torch.cuda.synchronize()
    def run(n0, n):
        for t in range(n0, n0 + n):
            for s, g in enumerate(engs):
                g.step_device(acts.data_ptr() + (t * total + s * E) * A * 4)
        for g in engs:
            g.sync()
    run(0, 100)
    t0 = time.perf_counter()
    run(100, K - 100)
    dt = (time.perf_counter() - t0) / (K - 100)
    print("%s total %d envs x %d agents as %d engine(s): %.1f us/step = %.2fM obs/s, faults %s" % (scenario, total, A, S, dt * 1e6, total * A / dt / 1e6, [g.faults() for g in engs]))
    for g in engs:
        g.close()
