#!/usr/bin/env python
"""BASELINE config 5: the eight Megaverse scenarios mixed, envs sharded over the GPUs of one node, optional NCCL gather of the
final observation tensor to every rank (env i runs scenario i % 8, so each GPU holds all eight).

    python tools/megaverse8_mixed.py --envs_per_gpu 1024 --steps 300                         # one GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/megaverse8_mixed.py --gather

One engine per scenario per rank, each on its own stream: their kernels overlap on the GPU.  Whole-job throughput = sum of
agent observations over ranks / max-over-ranks device time."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("BOXOBAN_LEVELS", os.path.join(ROOT, "tests", "golden", "boxoban"))

from megaverse_b200 import capi, sharding  # noqa: E402
from megaverse_b200.megaverse_env import MEGAVERSE8  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs_per_gpu", type=int, default=1024)
    ap.add_argument("--agents", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--gather", action="store_true", help="all-gather the obs tensor over NCCL after every step")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    begin, end = sharding.shard_range(a.envs_per_gpu * world, world, rank)
    per = (end - begin) // len(MEGAVERSE8)
    engines = []
    for k, scenario in enumerate(MEGAVERSE8):
        g = capi.Engine(scenario, per, a.agents, 128, 72, num_threads=2, device=local)
        for e in range(per):
            g.seed_env(e, 42 + begin + e * len(MEGAVERSE8) + k)  # global env i = begin + e*8 + k runs scenario k
        g.reset()
        engines.append(g)
    n_local = per * len(MEGAVERSE8) * a.agents
    rng = np.random.default_rng(1 + rank)
    masks = torch.from_numpy((1 << rng.integers(0, 11, size=(64, n_local))).astype(np.int32)).cuda()
    obs_local = [torch.as_tensor(g.device_array("obs"), device="cuda") for g in engines]
    gathered = torch.empty((world, n_local, 72, 128, 4), dtype=torch.uint8, device="cuda") if (a.gather and world > 1) else None
    stage = torch.empty((n_local, 72, 128, 4), dtype=torch.uint8, device="cuda") if gathered is not None else None

    def step(t):
        for k, g in enumerate(engines):
            g.step_device(masks.data_ptr() + ((t % 64) * n_local + k * per * a.agents) * 4)
        if gathered is not None:
            for g in engines:
                g.sync()
            torch.cat(obs_local, out=stage)
            dist.all_gather_into_tensor(gathered.view(world * n_local, 72, 128, 4), stage)

    for t in range(20):
        step(t)
    for g in engines:
        g.sync()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for t in range(a.steps):
        step(20 + t)
    for g in engines:
        g.sync()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    ups, max_ms, total = sharding.aggregate_throughput(n_local * a.steps, ms, dist if world > 1 else None)
    faults = [g.faults() for g in engines]
    if rank == 0:
        print("Megaverse-8 mixed: %d GPUs x %d envs (%d per scenario) x %d agents, %d steps, gather=%s: %.2fM obs/s whole job (%.3f ms/step), faults %s"
              % (world, per * 8, per, a.agents, a.steps, bool(gathered is not None), ups / 1e6, max_ms / a.steps, faults))
    for g in engines:
        g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
