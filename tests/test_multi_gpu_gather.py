"""BASELINE config 5 on two GPUs (needs `gpurun --gpus 2`; skipped on a single-GPU box): the eight scenarios mixed, one engine per scenario
rasterising into slices of one contiguous tensor, NCCL all-gather ordered by events.  Every rank's block of the gathered tensor must equal
what that rank rendered (checksums exchanged over NCCL), with no fault bit."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_gather_delivers_every_ranks_frames(built):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tools", "megaverse8_mixed.py"), "--gather", "--steps", "12", "--warmup", "3"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    rec = json.loads([ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1])
    assert rec["gathered_blocks_match_their_ranks"] is True, rec
    assert rec["faults"] == 0 and rec["value"] > 0, rec
