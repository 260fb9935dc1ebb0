#!/usr/bin/env python
"""The reference's own throughput harness (src/apps/megaverse_test_app.cpp:40-175,196-280) over this engine, same flags:

    python tools/megaverse_test_app.py --scenario Collect --num_envs 64 --num_agents 4 --performance_test

Random single-bit actions from a mt19937-like stream, env i seeded 42 + i, runs until 400 000 agent frames (with
--performance_test) and prints the FPS line of the reference ("Avg FPS" over agent frames, obs delivered to host memory)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megaverse_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenario", default="ObstaclesHard")
    ap.add_argument("--num_agents", type=int, default=2)
    ap.add_argument("--num_envs", type=int, default=64)
    ap.add_argument("--num_simulation_threads", type=int, default=1)
    ap.add_argument("--performance_test", action="store_true")
    ap.add_argument("--hires", action="store_true")
    ap.add_argument("--use_opengl", action="store_true", help="accepted for compatibility; there is one renderer")
    ap.add_argument("--device_resident", action="store_true", help="leave the observations in HBM (no D2H), action masks uploaded once")
    a = ap.parse_args()
    W, H = (800, 448) if a.hires else (128, 72)  # the reference uses 800x450; tiles need a height that is a multiple of 4
    max_frames = 400_000 if a.performance_test else 2_000_000
    E, A = a.num_envs, a.num_agents
    eng = capi.Engine(a.scenario, E, A, W, H, num_threads=max(1, a.num_simulation_threads))
    for e in range(E):
        eng.seed_env(e, 42 + e)  # :250-254
    eng.reset()
    rng = np.random.default_rng(42)
    steps = (max_frames + E * A - 1) // (E * A)
    print("Rendering resolution is [%dx%d] per agent; %d envs x %d agents, %d steps" % (W, H, E, A, steps))
    if a.device_resident:
        import torch

        masks = torch.from_numpy((1 << rng.integers(0, 11, size=(min(steps, 4096), E * A))).astype(np.int32)).cuda()
        t0 = time.perf_counter()
        for t in range(steps):
            eng.step_device(masks.data_ptr() + (t % masks.shape[0]) * E * A * 4)
        eng.sync()
    else:
        t0 = time.perf_counter()
        for t in range(steps):
            eng.step((1 << rng.integers(0, 11, size=E * A)).astype(np.int32))  # :140-147
    dt = time.perf_counter() - t0
    frames = steps * E * A
    print("Avg FPS: %.1f (agent observations per second; %d frames in %.2f s), faults %d" % (frames / dt, frames, dt, eng.faults()))
    published = {("empty", 64, 1): 75000, ("collect", 64, 1): 27000}.get((a.scenario.lower(), E, A))  # README.md:243-247, 10-core i9, its own GPU renderer
    if published and a.performance_test and not a.hires:
        print("reference README figure for this command line: approximately %d FPS" % published)
    eng.close()


if __name__ == "__main__":
    main()
