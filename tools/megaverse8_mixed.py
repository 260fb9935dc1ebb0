#!/usr/bin/env python
"""BASELINE config 5 on its own: the eight Megaverse scenarios mixed, envs sharded over the GPUs of one node, optional NCCL gather of the
observation tensor to every rank (global env i runs scenario i % 8, so each GPU holds all eight).  The measurement itself
is bench.py's (`measure_mixed`): one engine per scenario per rank, each on its own stream, all rasterising into slices of one
contiguous tensor (mv_set_obs_buffer); the gather is ordered by events, without host synchronisation.

    python tools/megaverse8_mixed.py --steps 300                                                             # one GPU
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/megaverse8_mixed.py --gather
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--gather", action="store_true", help="all-gather the obs tensor over NCCL after every step")
    ap.add_argument("--overlap", action="store_true", help="with --gather: double-buffer the obs tensor so that the gather of step t runs under step t+1")
    ap.add_argument("--grid-shares", action="store_true", help="every engine's raster grid takes 1/8 of the GPU's CTA slots (option raster_grid)")
    a = ap.parse_args()
    import torch
    import torch.distributed as dist

    rank, local, world = bench.dist_env()
    torch.cuda.set_device(local)
    bench.bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    hz = bench.Harness(torch, dist, world, local)
    rec = bench.measure_mixed(hz, a.steps, a.warmup, rank, os.cpu_count() or 1, gather=a.gather, overlap_gather=a.overlap, grid_share=a.grid_shares)
    if rank == 0:
        print(json.dumps(rec))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
