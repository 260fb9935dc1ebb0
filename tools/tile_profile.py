#!/usr/bin/env python
"""Per-tile cost distribution of mvr::tileKernel (clock64 per tile, mv_debug_tile_profile)."""
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
g.set_option("overlap", 0)
rng = np.random.default_rng(1)
for t in range(200):
    g.step((1 << rng.integers(0, 11, size=E * A)).astype(np.int32))
g.tile_profile(True, False)
g.step((1 << rng.integers(0, 11, size=E * A)).astype(np.int32))
p = g.tile_profile(True, True).astype(np.int64)
cyc, nov, nsm, nbg = p[..., 0].ravel(), (p[..., 1] & 0xfff).ravel(), (p[..., 2] & 0xfff).ravel(), p[..., 3].ravel()
tA, tB = (p[..., 1] >> 12).ravel(), (p[..., 2] >> 12).ravel()
print("phases (cycles, mean / p50): set-up %.0f / %.0f; lists + coverage + depth %.0f / %.0f; resolve+shade+store %.0f / %.0f" % (
    tA.mean(), np.median(tA), tB.mean(), np.median(tB), (cyc - tA - tB).mean(), np.median(cyc - tA - tB)))
print("%s E=%d A=%d kernel ms %s" % (scenario, E, A, g.last_kernel_ms()))
print("tiles %d; cycles/tile mean %.0f p50 %.0f p90 %.0f p99 %.0f max %d; sum %.1f Mcycles" % (cyc.size, cyc.mean(), np.median(cyc), np.percentile(cyc, 90), np.percentile(cyc, 99), cyc.max(), cyc.sum() / 1e6))
print("overlapping tris/tile mean %.1f p90 %.0f max %d; small mean %.1f; big mean %.1f max %d" % (nov.mean(), np.percentile(nov, 90), nov.max(), nsm.mean(), nbg.mean(), nbg.max()))
A_ = np.stack([np.ones_like(cyc), nsm, nbg], axis=1).astype(np.float64)
coef, *_ = np.linalg.lstsq(A_, cyc.astype(np.float64), rcond=None)
print("least squares: cycles ~ %.0f + %.1f * small + %.1f * big" % tuple(coef))
view_cyc = p[..., 0].sum(axis=1)
print("per-view cycles: mean %.0f max %.0f (x%.2f)" % (view_cyc.mean(), view_cyc.max(), view_cyc.max() / view_cyc.mean()))
g.close()
