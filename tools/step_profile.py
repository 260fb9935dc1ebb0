#!/usr/bin/env python
"""Per-phase latency of mvk::stepKernel (clock64 stamps per env, mv_debug_step_profile): where the per-env serial chain goes."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megaverse_b200 import capi

scenario = sys.argv[1] if len(sys.argv) > 1 else "TowerBuilding"
E = int(sys.argv[2]) if len(sys.argv) > 2 else 256
A = int(sys.argv[3]) if len(sys.argv) > 3 else 1
g = capi.Engine(scenario, E, A, 128, 72, num_threads=8)
for e in range(E):
    g.seed_env(e, 42 + e)
g.reset()
rng = np.random.default_rng(1)
names = ["stage", "action", "candlist", "kcc", "xform", "scenario", "out/reset", "instances", "commit"]
g.set_option("overlap", int(os.environ.get("MV_OVERLAP", "1")))
g.step_profile(True, False)
acc = []
for t in range(300):
    g.step((1 << rng.integers(0, 11, size=E * A)).astype(np.int32))
    if t >= 100:
        acc.append(g.step_profile(True, True).astype(np.int64))
acc = np.stack(acc)  # [T, E, 16]
stamps = acc[:, :, :9]
d = np.diff(np.concatenate([np.zeros_like(stamps[:, :, :1]), stamps], axis=2), axis=2)
print("scenario %s E=%d A=%d; kernel ms %s" % (scenario, E, A, g.last_kernel_ms()))
print("total cycles per env-step: mean %.0f  p50 %.0f  p99 %.0f  max-per-step mean %.0f" % (stamps[:, :, 8].mean(), np.median(stamps[:, :, 8]), np.percentile(stamps[:, :, 8], 99), stamps[:, :, 8].max(axis=1).mean()))
for i, n in enumerate(names):
    print("%-10s mean %7.0f  p99 %7.0f  max %7.0f   (in the slowest env of each step: %7.0f)" % (n, d[:, :, i].mean(), np.percentile(d[:, :, i], 99), d[:, :, i].max(),
          np.mean([d[t, np.argmax(stamps[t, :, 8]), i] for t in range(d.shape[0])])))
if acc[:, :, 9].max() > 0:
    sw, rc, cs, cr = acc[:, :, 9], acc[:, :, 10], acc[:, :, 11], acc[:, :, 13]
    print("kcc counters: sweeps mean %.2f max %d; recovers mean %.2f max %d; cycles/sweep %.0f; cycles/recover %.0f" % (sw.mean(), sw.max(), rc.mean(), rc.max(), cs.sum() / max(1, sw.sum()), cr.sum() / max(1, rc.sum())))
    kc = d[:, :, 3]
    idx = np.unravel_index(np.argsort(kc, axis=None)[-5:], kc.shape)
    for t, e in zip(*idx):
        print("  slow kcc %6d cycles: sweeps %d (%d cyc) recovers %d (%d cyc) cands %d" % (kc[t, e], sw[t, e], cs[t, e], rc[t, e], cr[t, e], acc[t, e, 12]))
print("candidates: mean %.1f max %d" % (acc[:, :, 12].mean(), acc[:, :, 12].max()))
g.close()
