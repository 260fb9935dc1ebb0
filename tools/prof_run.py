#!/usr/bin/env python
"""Short device-resident run of one config for ncu: prof_run.py <scenario> <envs> <agents> [depth] [steps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from megaverse_b200 import capi

scenario, E, A = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
depth = len(sys.argv) > 4 and sys.argv[4] == "depth"
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 30
g = capi.Engine(scenario, E, A, 128, 72, num_threads=16, depth=depth)
for k, v in [kv.split("=") for kv in os.environ.get("MV_OPTS", "").split(",") if kv]:
    g.set_option(k, int(v))
for e in range(E):
    g.seed_env(e, 42 + e)
g.reset()
acts = torch.from_numpy((1 << np.random.default_rng(1).integers(0, 11, size=(steps, E * A))).astype(np.int32)).cuda()
torch.cuda.synchronize()
for t in range(steps):
    g.step_device(acts.data_ptr() + t * E * A * 4)
g.sync()
print("done", g.faults())
g.close()
