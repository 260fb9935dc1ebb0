#!/usr/bin/env python
"""bench.py -- agent observations per second of the step+render hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA engine through the C ABI)
  python bench.py --impl reference --gpus N --steps K ...  the reference arm: the CPU restatement (oracle/) on the box's
                                                           host cores -- the real reference cannot be built here
                                                           (Bullet 2.89 / Vulkan / EGL absent)

A "step" is one pass of the hot path over one batch: TowerBuilding, 256 envs x 1 agent per GPU, 128x72 RGBA obs
(BASELINE.json configs[1]); weak scaling (256 envs on every GPU, no data-path collective).

  value  whole-job obs/s with the action masks already resident in HBM and the obs tensor left in HBM
  e2e    the same metric through the public host-buffer call (mv_set_actions + mv_step): H2D actions and D2H
         obs/rewards/dones inside the timed region
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENARIO, ENVS_PER_GPU, AGENTS, W, H = "TowerBuilding", 256, 1, 128, 72
OBS_BYTES = W * H * 4
METRIC, UNIT = "agent obs/sec (whole box)", "obs/s"
# dram__bytes_read.sum + dram__bytes_write.sum of geomKernel + tileKernel per launch at 256 envs (profiles/r1c_summary.txt: 7.95 + 6.21
# + 2.83 MB); the 9.4 MB of observations themselves stay in the 126 MB L2 until the consumer reads them
NCU_DRAM_BYTES_PER_STEP = 16736000  # geomKernel (7.68 MB read + 6.26 MB written) + tileKernel (2.79 MB read) per launch pair, profiles/r1h_summary.txt


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region"""

    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def action_stream(steps, n, seed):
    """the reference harness' distribution: one uniformly random action bit per agent per step (megaverse_test_app.cpp:140-147)"""
    rng = np.random.default_rng(seed)
    return (1 << rng.integers(0, 11, size=(steps, n))).astype(np.int32)


def run_cpu(steps, warmup, threads, envs):
    """times the oracle (CPU restatement) on the host cores; returns obs/s"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc

    o = orc.Oracle(SCENARIO, envs, AGENTS, W, H, threads=threads)
    for e in range(envs):
        o.seed_env(e, 42 + e)
    o.reset()
    acts = action_stream(steps + warmup, envs * AGENTS, 1)
    for t in range(warmup):
        o.step(acts[t])
    t0 = time.perf_counter()
    for t in range(warmup, warmup + steps):
        o.step(acts[t])
    dt = time.perf_counter() - t0
    o.close()
    return envs * AGENTS * steps / dt, dt


def run_reference_env_library(envs, threads, seconds=4.0):
    """simulation-only rate of the REFERENCE's own env library where its build travelled with the snapshot (oracle/_ref/pyref: the
    reference's pybind module + env.cpp / agent.cpp / character controller / scenarios compiled in place on the Bullet stand-in, null
    renderer -- DESIGN.md section 6).  No rendering: the reference renders on the GPU.  Returns None when the module is not there."""
    try:
        d = os.path.join(ROOT, "oracle", "_ref", "pyref")
        if not os.path.isdir(d):
            return None
        sys.path.insert(0, d)
        import megaverse as ref_ext

        ref_ext.set_megaverse_log_level(2)
        g = ref_ext.MegaverseGym(SCENARIO, W, H, envs, AGENTS, threads, True, {})
        g.seed(42)
        g.reset()
        rng = np.random.default_rng(1)
        heads = rng.integers(0, [3, 3, 3, 2, 2, 3], size=(64, envs * AGENTS, 6)).tolist()

        def one(t):
            row = heads[t % 64]
            for e in range(envs):
                for a in range(AGENTS):
                    g.set_actions(e, a, row[e * AGENTS + a])
            g.step()

        for t in range(10):
            one(t)
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            one(n)
            n += 1
        dt = time.perf_counter() - t0
        g.close()
        return {"value": envs * AGENTS * n / dt, "unit": "agent steps/s (simulation only, no rendering)", "threads": threads, "steps": n,
                "note": "reference env library compiled in place on the Bullet stand-in (analytic narrow phase), driven through its own pybind module"}
    except Exception as ex:  # never let the context figure break the bench line
        return {"unavailable": str(ex)[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank, local_rank, world = dist_env()
    K, Wm = args.steps, max(args.warmup, 3)
    E = args.envs_per_gpu
    N = E * AGENTS
    cores = os.cpu_count() or 1
    config = {"workload": "TowerBuilding num_envs=%d num_agents_per_env=%d %dx%d RGBA8 per GPU, random one-bit actions, resets included" % (E, AGENTS, W, H),
              "envs_per_gpu": E, "agents_per_env": AGENTS, "resolution": [W, H], "parallelism": "env-sharded x%d (no data-path collective)" % max(world, 1),
              "l2_policy": "L2 flushed (256 MB written) before EVERY timed step, untimed: each step is timed on its own with CUDA events on the "
                           "engine stream and the K step times are summed; value_l2_warm / e2e_l2_warm are the same loops run back to back "
                           "(steady-state rollout, per-step working set ~%.0f MB stays in the 126 MB L2)" % (N * OBS_BYTES / 1e6 + 12)}

    if args.impl == "reference":
        # the reference's own CPU path cannot be built here; the port (oracle) stands in.  Rank 0 only.
        if rank != 0:
            return
        k = min(K, 4000)  # ~15 s at 256 envs on a 128-thread host: each "step" is one pass over the same 256-env batch
        v, dt = run_cpu(k, min(Wm, 10), cores, E)
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": k, "warmup": min(Wm, 10),
                "ms_per_step": dt / k * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": "%d envs x %d steps (step + software render), all host threads" % (E, k)},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from megaverse_b200 import _build, capi

    _build.build_all()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    eng = capi.Engine(SCENARIO, E, AGENTS, W, H, num_threads=min(8, max(1, cores // max(world, 1))), device=local_rank)
    from megaverse_b200 import sharding

    begin, end = sharding.shard_range(E * world, world, rank)  # weak scaling: E envs on every rank
    for e, seed in enumerate(sharding.env_seeds(begin, end)):
        eng.seed_env(e, seed)  # global env i is seeded 42 + i (megaverse_test_app.cpp:250-254)
    eng.reset()
    stream = torch.cuda.ExternalStream(eng.stream(), device=local_rank)
    acts_host = action_stream(K + Wm, N, 1 + rank)
    acts_dev = torch.from_numpy(acts_host).cuda()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, base):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        for t in range(steps):
            fn(base + t)
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        _, ms, _ = sharding.aggregate_throughput(1, ms, dist if world > 1 else None)  # max over ranks of the device time
        return ms

    step_bytes = N * 4
    ptr0 = acts_dev.data_ptr()

    def dev_step(t):
        eng.step_device(ptr0 + t * step_bytes)  # asynchronous: the timed region ends with barrier() = device synchronize

    def host_step(t):
        eng.step(acts_host[t % len(acts_host)])

    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def timed_flushed(fn, steps, base):
        """K steps, each preceded by an (untimed) L2 flush; per-step CUDA events on the engine stream, summed; max over ranks"""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        barrier()
        for t in range(steps):
            with torch.cuda.stream(stream):
                flush.zero_()
            evs[t][0].record(stream)
            fn(base + t)
            evs[t][1].record(stream)
        eng.sync()
        barrier()
        ms = float(sum(a.elapsed_time(b) for a, b in evs))
        _, ms, _ = sharding.aggregate_throughput(1, ms, dist if world > 1 else None)
        return ms

    def timed_host_flushed(fn, steps, base):
        """the host-buffer call blocks until the results are in host memory: wall clock around each call, L2 flushed (and the
        flush waited for) before it"""
        total = 0.0
        barrier()
        for t in range(steps):
            with torch.cuda.stream(stream):
                flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn(base + t)
            total += time.perf_counter() - t0
        barrier()
        _, ms, _ = sharding.aggregate_throughput(1, total * 1e3, dist if world > 1 else None)
        return ms

    # ---- device-resident value (action masks resident in HBM, obs left in HBM)
    for t in range(Wm):
        dev_step(t)
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = eng.kernel_launches()
    ms = timed_flushed(dev_step, K, Wm)
    launches = eng.kernel_launches() - l0
    ms_warm = timed(dev_step, K, Wm)
    eng.sync()
    clocks = sampler.stop()
    value = N * world * K / (ms / 1e3)
    value_warm = N * world * K / (ms_warm / 1e3)

    # ---- end-to-end through host buffers
    Ke = max(50, min(K, 500))
    for t in range(3):
        host_step(t)
    ms_e = timed_host_flushed(host_step, Ke, Wm)
    e2e = N * world * Ke / (ms_e / 1e3)
    ms_e_warm = timed(host_step, Ke, Wm)
    e2e_warm = N * world * Ke / (ms_e_warm / 1e3)

    # ---- roofline of the dominant kernel (rasteriser): CUDA events around the kernel on its own stream, L2 flushed before
    peak, peak_src = measured_peaks()
    eng.set_option("overlap", 0)  # kernels back to back so that each can be timed on its own
    ras, stp = [], []
    for t in range(60):
        with torch.cuda.stream(stream):
            flush.zero_()
        dev_step(Wm + (t % K))
        eng.sync()
        s_ms, r_ms = eng.last_kernel_ms()
        stp.append(s_ms); ras.append(r_ms)
    ras_ms, stp_ms = float(np.mean(ras[10:])), float(np.mean(stp[10:]))
    achieved = N * OBS_BYTES / (ras_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "kernel": "mvr::geomKernel + mvr::tileKernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_STEP if E == ENVS_PER_GPU else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": N * OBS_BYTES, "kernel_ms": ras_ms, "step_kernel_ms": stp_ms,
                "note": "obs-write bytes / rasteriser duration; the kernels are latency / issue bound on short dependent chains (3.8 triangles per 32x4 tile), not HBM bound: DESIGN.md section 10"}
    faults = eng.faults()
    eng.close()

    # ---- the same host-delivered metric with a double-buffered consumer (mv_step_begin / mv_step_end): G engines of E/G envs each,
    # a group's next step is begun as soon as its previous result has been read, so one group's device->host copy (copy engine)
    # overlaps the other groups' kernels -- the way Sample Factory drives two env groups per worker.  Reported beside e2e.value.
    double_buffered = None
    G = 4
    if E % G == 0 and E // G >= 8:
        seeds = list(sharding.env_seeds(begin, end))
        Eg = E // G
        groups = []
        for g in range(G):
            eg = capi.Engine(SCENARIO, Eg, AGENTS, W, H, num_threads=max(1, min(8, cores // max(world, 1)) // G), device=local_rank)
            eg.set_option("zero_copy", 0)
            for e in range(Eg):
                eg.seed_env(e, seeds[g * Eg + e])
            eg.reset()
            groups.append(eg)
        Ng = Eg * AGENTS

        def db_loop(t0, n):
            sink = 0
            for g, eg in enumerate(groups):
                eg.step_begin(acts_host[t0 % len(acts_host)][g * Ng:(g + 1) * Ng])
            for t in range(t0 + 1, t0 + n):
                row = acts_host[t % len(acts_host)]
                for g, eg in enumerate(groups):
                    eg.step_end()
                    sink += int(eg.obs()[0, 0, 0, 0]) + int(eg.dones()[0])  # the consumer touches the delivered result
                    eg.step_begin(row[g * Ng:(g + 1) * Ng])
            for eg in groups:
                eg.step_end()
            return sink

        db_loop(0, 20)
        barrier()
        t0 = time.perf_counter()
        db_loop(20, Ke)
        ms_db = (time.perf_counter() - t0) * 1e3
        barrier()
        _, ms_db, _ = sharding.aggregate_throughput(1, ms_db, dist if world > 1 else None)
        fdb = sum(eg.faults() for eg in groups)
        for eg in groups:
            eg.close()
        double_buffered = {"value": N * world * Ke / (ms_db / 1e3), "unit": UNIT, "groups": G, "envs_per_group": Eg, "ms_per_step": ms_db / Ke, "faults": int(fdb),
                           "note": "wall clock over %d steps of all groups; obs via the copy engine (zero_copy=0); python consumer" % Ke}

    cpu_baseline = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        ksample = 300  # then scaled to ~15 s of CPU work
        v, dt = run_cpu(ksample, 3, cores, E)
        if dt < 12.0:
            ksample = int(min(20000, ksample * 15.0 / max(dt, 1e-3)))
            v, dt = run_cpu(ksample, 3, cores, E)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "%d envs x %d steps of the same workload on all %d host threads (%.1f s)" % (E, ksample, cores, dt),
                        "reference_env_library": run_reference_env_library(E, min(cores, 16))}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms / K, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "value_l2_warm": value_warm, "ms_per_step_l2_warm": ms_warm / K,
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": N * 4, "d2h_bytes_per_step": N * OBS_BYTES + N * 8 + E, "steps": Ke, "ms_per_step": ms_e / Ke,
                        "value_l2_warm": e2e_warm, "double_buffered": double_buffered},
                "clocks": clocks, "gpu_launches": int(launches), "faults": int(faults)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
