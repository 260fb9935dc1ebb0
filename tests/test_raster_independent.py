"""An INDEPENDENT restatement of the fixed-function rasterisation rules against the oracle's rasteriser (CPU only).

The reference draws with Vulkan (V4R: vulkan_state.cpp:588-606 -- back-face culling with counter-clockwise fronts, depth test
LESS_OR_EQUAL on a D32 attachment, one sample per pixel) and its tests hold no pixel vector, so the oracle's triangle rasteriser
(oracle/orc_raster.hpp) could so far only be compared with itself.  The rules it restates are not reference code but the Vulkan
specification's: view-volume clipping 0 <= z_c <= w_c, the viewport transform, fixed-point vertex positions (8 sub-pixel bits),
one sample at the pixel centre, the top-left rule for samples exactly on an edge, depth interpolated linearly in window space,
perspective-correct interpolation of the varyings (here: clip-space w, which V4R writes out as its depth image).

This file implements those rules a second time, in a different form and in different arithmetic:

  * coverage: exact integers, a sample on an edge is decided by displacing it infinitesimally to the right and, second order, down
    (lexicographic sign of (E, dE/dx, dE/dy)) -- no top-left classification of edges, no bias constants;
  * depth and w: float64 barycentrics from the exact integer edge values;
  * hidden-surface removal: per pixel over all covering triangles, nearest depth, the later draw on ties;
  * clipping: Sutherland-Hodgman in float64 on the clip coordinates.

Only the vertex stage is shared knowledge (float32 arithmetic in the order of uber.vert / Magnum, pinned elsewhere against the real
shader text and the real Magnum): it is recomputed here in numpy float32 so that both sides snap the same window coordinates.

Compared: the oracle's depth image (view-space w of the visible fragment, 0 = nothing drawn) -- its zero pattern IS the coverage
mask.  Scenes without clipping must agree in every pixel (coverage exactly, depth to float32 rounding); scenes cut by the near
plane (the clipper creates new vertices, whose float32 vs float64 positions can snap one sub-pixel apart) in all but a handful.
"""
import numpy as np
import pytest

import orc

W, H = 128, 72
F32 = np.float32


def _projection():
    aspect = F32(W) / F32(H)
    half_tan = F32(np.tan(np.float64(F32(100.0) * F32(0.01745329251994329576923690768489)) / 2.0))  # correctly rounded tan, as the oracle's crtan
    near, far = F32(0.01), F32(120.0)
    return F32(1.0) / half_tan, -aspect / half_tan, far / (near - far), far * near / (near - far)


def _mesh(kind):
    vtx, idx = orc.mesh(kind)
    return vtx.view(np.float32).reshape(-1, 6)[:, :3].copy(), idx.reshape(-1, 3).astype(np.int64)


def _vertex_stage(view16, model16, verts):
    """clip-space positions, float32, in the operation order of the vertex stage (uber.vert:53-110 on Magnum's column-major matrices)"""
    v = view16.reshape(4, 4)   # v[col][row]
    m = model16.reshape(4, 4)
    mv = np.zeros((4, 4), dtype=F32)
    for col in range(4):
        for row in range(4):
            acc = F32(0.0)
            for pos in range(4):
                acc = F32(acc + F32(v[pos][row] * m[col][pos]))
            mv[col][row] = acc
    p00, p11, p22, p32 = _projection()
    out = np.zeros((len(verts), 4), dtype=F32)
    for i, p in enumerate(verts):
        cam = []
        for row in range(3):
            acc = F32(0.0)
            acc = F32(acc + F32(mv[0][row] * p[0])); acc = F32(acc + F32(mv[1][row] * p[1])); acc = F32(acc + F32(mv[2][row] * p[2])); acc = F32(acc + F32(mv[3][row] * F32(1.0)))
            cam.append(acc)
        out[i] = (F32(cam[0] * p00), F32(cam[1] * p11), F32(F32(cam[2] * p22) + p32), F32(-cam[2]))
    return out


def _snap(clip):
    """viewport transform + fixed point, the window coordinates a rasteriser with 8 sub-pixel bits works on (float32 like the oracle)"""
    r = F32(1.0) / clip[3]
    hw, hh = F32(W) * F32(0.5), F32(H) * F32(0.5)
    x = F32(F32(F32(clip[0] * r) * hw) + hw)
    y = F32(F32(F32(clip[1] * r) * hh) + hh)
    return int(np.floor(np.float64(F32(F32(x * F32(256.0)) + F32(0.5))))), int(np.floor(np.float64(F32(F32(y * F32(256.0)) + F32(0.5))))), np.float64(F32(clip[2] * r)), np.float64(r)


def _clip_polygon(poly):
    """Sutherland-Hodgman against 0 <= z <= w in float64 (Vulkan 'primitive clipping', depth range zero-to-one)"""
    for plane in (0, 1):
        dist = (lambda c: c[2]) if plane == 0 else (lambda c: c[3] - c[2])
        out = []
        for i in range(len(poly)):
            a, b = poly[i], poly[(i + 1) % len(poly)]
            da, db = dist(a), dist(b)
            if da >= 0:
                out.append(a)
            if (da >= 0) != (db >= 0):
                t = da / (da - db)
                out.append(a + t * (b - a))
        poly = out
        if len(poly) < 3:
            return []
    return poly


STATS = {"ties": 0}  # samples that lay exactly on an edge of a front-facing triangle (the cases the tie rule decides)


def _independent_depth(view16, instances):
    """depth image (view-space w of the visible fragment, 0 = empty) by the rules of the specification; also returns whether any
    primitive needed clipping"""
    ys, xs = np.mgrid[0:H, 0:W]
    sx = (xs * 256 + 128).astype(np.int64)
    sy = (ys * 256 + 128).astype(np.int64)
    best_z = np.full((H, W), np.inf)
    best_w = np.zeros((H, W))
    clipped_any = False
    for row in instances:
        verts, tris = _mesh(int(row[0]))
        clip = _vertex_stage(view16, row[2:18].astype(F32), verts)
        for tri in tris:
            c = clip[tri]
            needs_clip = bool((c[:, 2] < 0).any() or (c[:, 3] - c[:, 2] < 0).any())
            if needs_clip:
                clipped_any = True
                poly = _clip_polygon([c[k].astype(np.float64) for k in range(3)])
                pieces = [(poly[0], poly[k], poly[k + 1]) for k in range(1, len(poly) - 1)]
            else:
                pieces = [(c[0], c[1], c[2])]
            for piece in pieces:
                sv = [_snap(np.asarray(v, dtype=F32)) for v in piece]
                (x0, y0, z0, r0), (x1, y1, z1, r1), (x2, y2, z2, r2) = sv
                area2 = (x1 - x0) * (y2 - y0) - (y1 - y0) * (x2 - x0)
                if area2 >= 0:
                    continue  # clockwise on a y-down screen = back face (front faces are counter-clockwise in y-up NDC), or degenerate
                inside = np.ones((H, W), dtype=bool)
                lam = []
                for (ax, ay), (bx, by) in (((x1, y1), (x2, y2)), ((x2, y2), (x0, y0)), ((x0, y0), (x1, y1))):
                    # edge function oriented so that the interior is positive; a sample ON the edge counts iff an infinitesimal step
                    # right (then down) enters the interior: sign of (E, dE/dx, dE/dy) in lexicographic order
                    dEdx, dEdy = (by - ay), -(bx - ax)
                    E = dEdx * (sx - ax) + dEdy * (sy - ay)
                    tie = dEdx > 0 or (dEdx == 0 and dEdy > 0)
                    inside &= (E > 0) | ((E == 0) & tie)
                    STATS["ties"] += int((E == 0).sum())
                    lam.append(E.astype(np.float64) / float(-area2))
                if not inside.any():
                    continue
                z = lam[0] * z0 + lam[1] * z1 + lam[2] * z2                 # depth: linear in window space
                w = 1.0 / (lam[0] * r0 + lam[1] * r1 + lam[2] * r2)        # perspective-correct w
                win = inside & (z <= 1.0) & (z <= best_z)                     # LESS_OR_EQUAL: the later draw replaces an equal depth
                best_z = np.where(win, z, best_z)
                best_w = np.where(win, w, best_w)
    return best_w, clipped_any


def _random_scene(rng, n, near_camera):
    """instances (mesh, colour, model matrix column-major) in front of a camera at the origin looking down -z with a random roll"""
    rows = []
    for _ in range(n):
        mesh = int(rng.choice([0, 0, 0, 1, 2, 3, 4]))
        s = rng.uniform(0.2, 1.5, size=3) * (3.0 if near_camera and mesh == 0 and rng.random() < 0.3 else 1.0)
        ang = rng.uniform(0, 2 * np.pi, size=3)
        cx, sx_ = np.cos(ang[0]), np.sin(ang[0]); cy, sy_ = np.cos(ang[1]), np.sin(ang[1]); cz, sz_ = np.cos(ang[2]), np.sin(ang[2])
        rx = np.array([[1, 0, 0], [0, cx, -sx_], [0, sx_, cx]]); ry = np.array([[cy, 0, sy_], [0, 1, 0], [-sy_, 0, cy]]); rz = np.array([[cz, -sz_, 0], [sz_, cz, 0], [0, 0, 1]])
        rot = rz @ ry @ rx @ np.diag(s)
        depth = rng.uniform(0.3, 3.0) if near_camera else rng.uniform(4.0, 14.0)
        t = np.array([rng.uniform(-1.2, 1.2) * depth, rng.uniform(-0.8, 0.8) * depth, -depth])
        m = np.eye(4); m[:3, :3] = rot; m[:3, 3] = t
        rows.append(np.concatenate([[mesh, rng.integers(0, 20)], m.T.reshape(-1)]))  # column-major
    rows.sort(key=lambda r: r[0])  # draw order: by mesh type, boxes first (what the product's instance lists guarantee)
    roll = rng.uniform(-0.3, 0.3)
    v = np.eye(4); v[:2, :2] = [[np.cos(roll), -np.sin(roll)], [np.sin(roll), np.cos(roll)]]
    return v.T.reshape(-1).astype(F32), np.array(rows, dtype=F32)


def _compare(view16, inst):
    _, ref = orc.render_instances(view16, inst, W, H, want_depth=True)
    mine, clipped = _independent_depth(view16, inst)
    cov_ref, cov_mine = ref > 0, mine > 0
    both = cov_ref & cov_mine
    rel = np.abs(ref[both].astype(np.float64) - mine[both]) / mine[both] if both.any() else np.zeros(0)
    return cov_ref, cov_mine, rel, clipped


@pytest.mark.parametrize("seed", range(8))
def test_unclipped_scenes_agree_in_every_pixel(seed):
    rng = np.random.default_rng(1000 + seed)
    view16, inst = _random_scene(rng, 6, near_camera=False)
    cov_ref, cov_mine, rel, clipped = _compare(view16, inst)
    assert not clipped
    assert cov_ref.sum() > 50, "the scene should cover some pixels"
    assert np.array_equal(cov_ref, cov_mine), "coverage differs in %d pixels" % int((cov_ref != cov_mine).sum())
    # same visible surface: w agrees to float32 rounding of the oracle's arithmetic (a different winner would be a different surface)
    assert rel.max() < 2e-5, "depth differs by %.3g" % rel.max()


@pytest.mark.parametrize("seed", range(6))
def test_scenes_cut_by_the_near_plane_agree(seed):
    rng = np.random.default_rng(2000 + seed)
    view16, inst = _random_scene(rng, 5, near_camera=True)
    cov_ref, cov_mine, rel, _ = _compare(view16, inst)
    assert cov_ref.sum() > 200
    mism = int((cov_ref != cov_mine).sum())
    assert mism <= 8, "coverage differs in %d pixels" % mism  # new vertices made by the clipper may snap one sub-pixel apart
    assert (rel > 1e-4).sum() <= 8, "visible surface differs in %d pixels" % int((rel > 1e-4).sum())


@pytest.mark.parametrize("seed", range(4))
def test_edges_through_pixel_centres(seed):
    """random scenes almost never put a sample exactly on an edge, so the tie rule needs scenes built for it: boxes facing the camera whose
    front-face corners project onto pixel centres (within the 1/512-pixel reach of the fixed-point snap) -- whole rows and columns of
    samples, and the diagonal the face's two triangles share, lie exactly on edges.  Coverage must still agree in every pixel."""
    rng = np.random.default_rng(3000 + seed)
    p00, p11, _, _ = _projection()
    rows = []
    for k in range(5):
        z_front = -rng.uniform(3.0, 9.0)
        half_z = rng.uniform(0.2, 0.6)
        px0, py0 = int(rng.integers(4, 90)), int(rng.integers(4, 50))
        wpx = int(rng.integers(6, 30)); hpx = wpx if k % 2 == 0 else int(rng.integers(6, 20))  # squares: the shared diagonal hits pixel centres too
        x0, x1 = [((px + 0.5) - W / 2) / (W / 2) * (-z_front) / float(p00) for px in (px0, px0 + wpx)]
        y0, y1 = [((py + 0.5) - H / 2) / (H / 2) * (-z_front) / float(p11) for py in (py0, py0 + hpx)]
        m = np.eye(4)
        m[0, 0], m[1, 1], m[2, 2] = abs(x1 - x0) / 2, abs(y1 - y0) / 2, half_z
        m[:3, 3] = [(x0 + x1) / 2, (y0 + y1) / 2, z_front - half_z]
        rows.append(np.concatenate([[0, k], m.T.reshape(-1)]))
    view16 = np.eye(4, dtype=F32).reshape(-1)
    inst = np.array(rows, dtype=F32)
    STATS["ties"] = 0
    cov_ref, cov_mine, rel, clipped = _compare(view16, inst)
    assert not clipped
    assert STATS["ties"] > 100, "the scene was built to put samples on edges (%d)" % STATS["ties"]
    assert np.array_equal(cov_ref, cov_mine), "coverage differs in %d pixels" % int((cov_ref != cov_mine).sum())
    assert rel.max() < 2e-5


def test_shared_edges_are_drawn_exactly_once():
    """the tie rule's purpose: two triangles sharing an edge cover every sample on it exactly once -- a fan of thin triangles around a
    vertex placed exactly on a pixel centre, edges through many pixel centres; checked on the independent rule itself and through the
    oracle (a full-screen quad of two triangles plus axis-aligned boxes whose edges run through pixel centres leave no hole)"""
    cx, cy = 64 * 256 + 128, 36 * 256 + 128
    ys, xs = np.mgrid[0:H, 0:W]
    sx = (xs * 256 + 128).astype(np.int64); sy = (ys * 256 + 128).astype(np.int64)
    ring = [(cx + int(30 * 256 * np.cos(a)) // 128 * 128, cy + int(30 * 256 * np.sin(a)) // 128 * 128) for a in np.linspace(0, 2 * np.pi, 17)[:-1]]
    count = np.zeros((H, W), dtype=int)
    for k in range(16):
        (x1, y1), (x2, y2) = ring[k], ring[(k + 1) % 16]
        tri = [(cx, cy), (x2, y2), (x1, y1)]
        area2 = (tri[1][0] - tri[0][0]) * (tri[2][1] - tri[0][1]) - (tri[1][1] - tri[0][1]) * (tri[2][0] - tri[0][0])
        if area2 > 0:
            tri = [tri[0], tri[2], tri[1]]
        inside = np.ones((H, W), dtype=bool)
        for (ax, ay), (bx, by) in ((tri[1], tri[2]), (tri[2], tri[0]), (tri[0], tri[1])):
            dEdx, dEdy = (by - ay), -(bx - ax)
            E = dEdx * (sx - ax) + dEdy * (sy - ay)
            inside &= (E > 0) | ((E == 0) & (dEdx > 0 or (dEdx == 0 and dEdy > 0)))
        count += inside
    assert count.max() == 1, "a sample was covered twice"
    assert count[36, 64] == 1, "the shared vertex on a pixel centre belongs to exactly one triangle"
    interior = (sx - cx) ** 2 + (sy - cy) ** 2 < (25 * 256) ** 2
    assert (count[interior] == 1).all(), "a sample inside the fan was not covered"


def _tie_scene(seed):
    rng = np.random.default_rng(3000 + seed)
    p00, p11, _, _ = _projection()
    rows = []
    for k in range(5):
        z_front = -rng.uniform(3.0, 9.0)
        half_z = rng.uniform(0.2, 0.6)
        px0, py0 = int(rng.integers(4, 90)), int(rng.integers(4, 50))
        wpx = int(rng.integers(6, 30)); hpx = wpx if k % 2 == 0 else int(rng.integers(6, 20))
        x0, x1 = [((px + 0.5) - W / 2) / (W / 2) * (-z_front) / float(p00) for px in (px0, px0 + wpx)]
        y0, y1 = [((py + 0.5) - H / 2) / (H / 2) * (-z_front) / float(p11) for py in (py0, py0 + hpx)]
        m = np.eye(4)
        m[0, 0], m[1, 1], m[2, 2] = abs(x1 - x0) / 2, abs(y1 - y0) / 2, half_z
        m[:3, 3] = [(x0 + x1) / 2, (y0 + y1) / 2, z_front - half_z]
        rows.append(np.concatenate([[0, k], m.T.reshape(-1)]))
    return np.eye(4, dtype=F32).reshape(-1), np.array(rows, dtype=F32)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,seed", [("far", 0), ("far", 1), ("ties", 0), ("ties", 1), ("near", 0)])
def test_cuda_rasteriser_against_the_independent_rules(kind, seed):
    """the CUDA kernel itself (mv_debug_render_instances: the product's viewKernel on one view) against this file's restatement of the
    specification -- no oracle in between"""
    from megaverse_b200 import capi

    if kind == "ties":
        view16, inst = _tie_scene(seed)
    else:
        rng = np.random.default_rng((1000 if kind == "far" else 2000) + seed)
        view16, inst = _random_scene(rng, 6 if kind == "far" else 5, near_camera=(kind == "near"))
    _, dev = capi.render_instances(view16, inst, W, H, want_depth=True)
    mine, _ = _independent_depth(view16, inst)
    cov_dev, cov_mine = dev > 0, mine > 0
    mism = int((cov_dev != cov_mine).sum())
    assert mism <= (8 if kind == "near" else 0), "coverage differs in %d pixels" % mism
    both = cov_dev & cov_mine
    rel = np.abs(dev[both].astype(np.float64) - mine[both]) / mine[both]
    assert (rel > (1e-4 if kind == "near" else 2e-5)).sum() <= (8 if kind == "near" else 0)
