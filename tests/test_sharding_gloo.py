"""The N>1 path on CPU: env-range sharding, per-env seeds, view offsets and the whole-job throughput reduction, exercised with
two real processes over the gloo backend (rendezvous on 127.0.0.1)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total_envs, agents, out_dir):
    import torch.distributed as dist

    from megaverse_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    begin, end = sharding.shard_range(total_envs, world, rank)
    seeds = sharding.env_seeds(begin, end)
    # pretend device times: rank 1 is the slow one
    local_ms = 10.0 + 5.0 * rank
    units = (end - begin) * agents * 100  # 100 steps
    ups, max_ms, total = sharding.aggregate_throughput(units, local_ms, dist)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array([begin, end, seeds[0], seeds[-1], sharding.view_offset(begin, agents), ups, max_ms, total], dtype=np.float64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_sharding_over_gloo(tmp_path):
    import torch.multiprocessing as mp

    world, total_envs, agents = 2, 513, 4  # odd total: shard sizes differ by one
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total_envs, agents, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(os.path.join(str(tmp_path), "r%d.npy" % k)) for k in range(world)]
    assert (r[0][0], r[0][1], r[1][0], r[1][1]) == (0, 257, 257, 513)  # contiguous, disjoint, complete
    assert r[0][2] == 42 and r[0][3] == 42 + 256 and r[1][2] == 42 + 257 and r[1][3] == 42 + 512  # seed of env i is 42 + i
    assert r[0][4] == 0 and r[1][4] == 257 * agents
    want_total = 513 * agents * 100
    for k in range(world):
        assert r[k][7] == want_total and r[k][6] == 15.0  # sum of units, max over ranks of the time
        assert abs(r[k][5] - want_total / 0.015) < 1e-6


def test_shard_ranges_cover_everything():
    from megaverse_b200 import sharding

    for total in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)
