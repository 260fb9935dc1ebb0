#!/usr/bin/env python
"""Host-delivered throughput with a double-buffered consumer: G engines of total/G envs each; engine g's step k+1 is begun as soon
as its step k has been consumed, so one group's device->host copy overlaps the other groups' kernels (mv_step_begin / mv_step_end)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megaverse_b200 import capi

scenario = sys.argv[1] if len(sys.argv) > 1 else "TowerBuilding"
total = int(sys.argv[2]) if len(sys.argv) > 2 else 256
A = int(sys.argv[3]) if len(sys.argv) > 3 else 1
K = 1500
rng = np.random.default_rng(1)
for G, zero_copy in ((1, 1), (1, 0), (2, 1), (2, 0), (4, 0), (2, 0)):
    E = total // G
    engs = []
    for g in range(G):
        e = capi.Engine(scenario, E, A, 128, 72, num_threads=max(1, 8 // G))
        e.set_option("zero_copy", zero_copy)
        for i in range(E):
            e.seed_env(i, 42 + g * E + i)
        e.reset()
        engs.append(e)
    acts = (1 << rng.integers(0, 11, size=(K + 200, G, E * A))).astype(np.int32)
    sink = 0
    def loop(t0, n):
        global sink
        for g, e in enumerate(engs):
            e.step_begin(acts[t0, g])
        for t in range(t0 + 1, t0 + n):
            for g, e in enumerate(engs):
                e.step_end()
                sink += int(e.obs()[0, 0, 0, 0]) + int(e.dones()[0])  # the consumer touches the result
                e.step_begin(acts[t, g])
        for e in engs:
            e.step_end()
    loop(0, 100)
    t0 = time.perf_counter()
    loop(100, K)
    dt = time.perf_counter() - t0
    print(f"{scenario} {total}x{A}: groups={G} zero_copy={zero_copy}  {dt / K * 1e6:7.1f} us per {total * A} obs  {total * A * K / dt / 1e6:.3f} M obs/s (host-delivered)", flush=True)
    for e in engs:
        e.close()
