#!/usr/bin/env python
"""Capacity check: every scenario at several agent counts for a long random rollout; prints the sticky fault mask (0 = clean)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from megaverse_b200 import capi

os.environ.setdefault("BOXOBAN_LEVELS", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "boxoban"))
names = {1: "LEVEL_NOT_READY", 2: "TRI_OVERFLOW", 4: "GRID_RANGE", 8: "NAN", 16: "ENVELOPE", 32: "CAND_OVERFLOW"}
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for scenario in ["TowerBuilding", "ObstaclesHard", "ObstaclesEasy", "Collect", "Sokoban", "Rearrange", "HexExplore", "HexMemory"]:
    for E, A in [(256, 1), (64, 4), (32, 8)]:
        try:
            g = capi.Engine(scenario, E, A, 128, 72, num_threads=16)
            for e in range(E):
                g.seed_env(e, 1000 + e)
            g.reset()
        except capi.MegaverseError as ex:  # e.g. more agents than the scenario's start area holds
            print("%-14s E=%-4d A=%d: not runnable: %s" % (scenario, E, A, ex))
            continue
        rng = np.random.default_rng(7)
        heads = rng.integers(0, 2, size=(512, E * A, 11)) * (rng.random(size=(512, E * A, 11)) < 0.25)
        masks = np.zeros((512, E * A), dtype=np.int32)
        for b in range(1, 11):
            masks |= (heads[:, :, b] << b).astype(np.int32)
        d = torch.from_numpy(masks).cuda()
        t0 = time.perf_counter()
        try:
            for t in range(steps):
                g.step_device(d.data_ptr() + (t % 512) * E * A * 4)
            g.sync()
        except capi.MegaverseError as ex:
            print("%-14s E=%-4d A=%d: stopped: %s" % (scenario, E, A, ex))
        f = g.faults()
        print("%-14s E=%-4d A=%d: %7.0f obs/s faults %d %s" % (scenario, E, A, steps * E * A / (time.perf_counter() - t0), f, [n for b, n in names.items() if f & b]))
        g.close()
