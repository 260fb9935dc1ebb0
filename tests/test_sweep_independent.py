"""An INDEPENDENT check of the analytic narrow phase that stands in for Bullet's convexSweepTest (Bullet 2.89 is
neither vendored nor installed -- the GPU box was probed as well, tools/probe_box.sh -- so the capsule sweep is the one part of the path
that cannot be pinned against the reference's own code).

Bullet answers a convex sweep by CONSERVATIVE ADVANCEMENT over a closest-distance query (btContinuousConvexCollision: advance by
(distance + allowedPenetration) / |relative velocity| until the gap closes to the tolerance).  This file restates exactly that scheme
in numpy on top of its own distance function -- segment-to-box distance by golden-section search over the capsule axis with the
closed-form point-to-box distance, nothing shared with oracle/orc_physics.hpp's feature enumeration -- and compares the time of impact
with the oracle's sweepNarrow on random sweeps against axis-aligned boxes, boxes turned about Y and other agents' capsules.

Two results:
  * same shapes (sharp-edged boxes): the two answers agree to the advancement's termination tolerance;
  * Bullet's btBoxShape carries a 0.04 collision margin (its corners and edges are rounded by that radius): the test MEASURES how much
    later a margin-rounded box is hit and bounds it by the margin's geometric maximum -- the size of the band in which a real Bullet
    build may differ from the analytic definition at box edges and corners (flat-face hits are identical).
"""
import ctypes as C

import numpy as np
import pytest

R, HALF, ALLOWED = 0.33, 0.525, 0.04  # agent.cpp:52-53, btDispatcherInfo::m_allowedCcdPenetration
GOLD = (np.sqrt(5.0) - 1.0) / 2.0


def point_box(p, h, margin=0.0):
    """distance from point p (box frame) to the box of half extents h; with margin: to the box shrunk by margin and rounded back"""
    q = np.abs(p) - (h - margin)
    outside = np.linalg.norm(np.maximum(q, 0.0))
    inside = min(float(q.max()), 0.0)
    return outside + inside - margin


def segment_box(c, h, margin=0.0):
    """distance from the vertical segment c +- HALF*y to the box: convex in the segment parameter -> golden section"""
    f = lambda s: point_box(c + np.array([0.0, s, 0.0]), h, margin)
    lo, hi = -HALF, HALF
    x1, x2 = hi - GOLD * (hi - lo), lo + GOLD * (hi - lo)
    f1, f2 = f(x1), f(x2)
    for _ in range(60):
        if f1 < f2:
            hi, x2, f2 = x2, x1, f1
            x1 = hi - GOLD * (hi - lo); f1 = f(x1)
        else:
            lo, x1, f1 = x1, x2, f2
            x2 = lo + GOLD * (hi - lo); f2 = f(x2)
    return min(f1, f2, f(-HALF), f(HALF))


def segment_segment_vertical(c):
    """distance between two vertical segments of half length HALF whose centres differ by c"""
    dy = max(abs(c[1]) - 2 * HALF, 0.0)
    return float(np.sqrt(c[0] ** 2 + c[2] ** 2 + dy ** 2))


def gap(col, p, margin=0.0):
    """surface distance between the agent capsule at p and the collider (negative = penetration)"""
    rel = p - col[1:4]
    if int(col[0]) == 1:
        return segment_segment_vertical(rel) - 2 * R
    yaw = float(col[7])
    if yaw != 0.0:  # box turned about Y: its local x axis is (cos yaw, 0, -sin yaw)
        ax, az = np.cos(yaw), -np.sin(yaw)
        rel = np.array([rel[0] * ax + rel[2] * az, rel[1], -rel[0] * az + rel[2] * ax])
    return segment_box(rel, col[4:7].astype(np.float64), margin) - R


def conservative_advancement(col, f, d, margin=0.0, tol=1e-5, max_iter=200):
    """first t in [0,1] at which the capsule penetrates by ALLOWED, Bullet's scheme; None = no hit"""
    length = float(np.linalg.norm(d))
    t = 0.0
    for _ in range(max_iter):
        g = gap(col, f + d * t, margin) + ALLOWED
        if g <= tol:
            return t
        t += g / max(length, 1e-12)  # the gap cannot close faster than the capsule moves
        if t > 1.0:
            return None
    return t


@pytest.fixture(scope="module")
def sweep_lib():
    import orc

    O = orc.lib()
    O.orc_sweep_case.argtypes = [C.c_void_p] * 5
    return O


def _cases(rng, n):
    for _ in range(n):
        kind = 0 if rng.random() < 0.8 else 1
        col = np.zeros(8, np.float32)
        col[0] = kind
        col[1:4] = rng.uniform(-2, 2, 3)
        col[4:7] = rng.uniform(0.3, 3.0, 3)
        col[7] = 0.0 if rng.random() < 0.5 else rng.uniform(-3.1, 3.1)
        a = rng.normal(size=3); a /= np.linalg.norm(a)
        reach = float(np.linalg.norm(col[4:7])) + 1.5
        f = (col[1:4] + a * rng.uniform(1.0, 1.4) * reach).astype(np.float32)  # starts clear of the collider
        to = (col[1:4] + rng.normal(size=3) * 0.5 * reach).astype(np.float32)
        yield col, np.ascontiguousarray(f), np.ascontiguousarray(to)


def test_sweep_agrees_with_independent_conservative_advancement(built, sweep_lib):
    rng = np.random.default_rng(2024)
    t_out, n_out = np.zeros(1, np.float32), np.zeros(3, np.float32)
    hits = misses = 0
    worst = 0.0
    for col, f, to in _cases(rng, 1500):
        d = (to - f).astype(np.float64)
        length = float(np.linalg.norm(d))
        if gap(col, f.astype(np.float64)) <= 0.05:
            continue  # the comparison is about first contact from outside
        hit = bool(sweep_lib.orc_sweep_case(col.ctypes.data, f.ctypes.data, to.ctypes.data, t_out.ctypes.data, n_out.ctypes.data))
        ca = conservative_advancement(col, f.astype(np.float64), d)
        if ca is None:
            # the advancement never closed the gap: the analytic sweep may only report a graze within the tolerance
            if hit:
                assert gap(col, f + d * float(t_out[0])) + ALLOWED < 2e-3, "analytic hit where the advancement finds no contact"
            misses += 1
            continue
        assert hit, "advancement reaches the tolerance surface at t = %.5f, the analytic sweep reports no hit" % ca
        err = abs(float(t_out[0]) - ca) * length  # distance along the sweep
        worst = max(worst, err)
        assert err < 2e-3, "time of impact differs by %.5f length units (analytic %.6f, advancement %.6f)" % (err, float(t_out[0]), ca)
        hits += 1
    assert hits > 300 and misses > 300, (hits, misses)
    print("conservative advancement vs analytic sweep: %d hits, worst |dt|*length = %.2e" % (hits, worst))


def test_effect_of_bullets_box_margin_is_bounded(built, sweep_lib):
    """how much later is a box with Bullet's 0.04 margin rounding hit?  Zero on faces; at edges at most (sqrt2 - 1) * 0.04 / cos(incidence),
    at corners (sqrt3 - 1) * 0.04 / cos(incidence): bounded here for sweeps that do not graze (incidence below ~60 degrees)"""
    rng = np.random.default_rng(7)
    t_out, n_out = np.zeros(1, np.float32), np.zeros(3, np.float32)
    deltas = []
    for col, f, to in _cases(rng, 900):
        if int(col[0]) != 0:
            continue
        d = (to - f).astype(np.float64)
        length = float(np.linalg.norm(d))
        if gap(col, f.astype(np.float64)) <= 0.05:
            continue
        if not sweep_lib.orc_sweep_case(col.ctypes.data, f.ctypes.data, to.ctypes.data, t_out.ctypes.data, n_out.ctypes.data):
            continue
        cos_inc = -float(np.dot(n_out.astype(np.float64), d)) / length
        if cos_inc < 0.5:
            continue
        ca = conservative_advancement(col, f.astype(np.float64), d, margin=0.04)
        if ca is None:
            continue  # the rounded box is missed where the sharp one is grazed
        delta = (ca - float(t_out[0])) * length
        assert -2e-3 < delta < (np.sqrt(3.0) - 1.0) * 0.04 / cos_inc + 2e-3, delta
        deltas.append(delta)
    deltas = np.array(deltas)
    assert len(deltas) > 150
    on_face = float((deltas < 1e-3).mean())
    print("box margin: %d sweeps, %.0f %% unchanged (face hits), mean +%.4f, max +%.4f length units" % (len(deltas), 100 * on_face, deltas.mean(), deltas.max()))
    assert on_face > 0.5
