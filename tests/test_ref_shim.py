"""Pins the oracle's restated math against the pieces of the REAL reference that compile in place (oracle/_ref/libmvref.so,
built by `make -C oracle ref` from /root/reference: vendored Magnum + the reference's util headers).  The .so travels to
the GPU box; where it is absent (no /root/reference and never built) these tests skip."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libmvref.so")


@pytest.fixture(scope="module")
def libs(built):
    import orc

    if not os.path.exists(REF):
        if os.path.isdir("/root/reference/src/3rdparty/magnum"):
            import subprocess

            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
        else:
            pytest.skip("oracle/_ref/libmvref.so not built and /root/reference absent")
    R, O = C.CDLL(REF), orc.lib()
    return R, O


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _rand_affine(rng):
    """scene-graph style matrices: rotation * non-uniform scale + translation"""
    ang = rng.uniform(-3, 3)
    ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
    c, s = np.cos(ang), np.sin(ang)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    Rm = np.eye(3) + s * K + (1 - c) * K @ K
    M = np.eye(4)
    M[:3, :3] = Rm @ np.diag(rng.uniform(0.1, 20, size=3))
    M[:3, 3] = rng.uniform(-30, 30, size=3)
    return _f(M.T.reshape(-1))  # column-major


def test_matrix_product_and_inverse_bit_exact(libs):
    """Magnum RectangularMatrix::operator* and Matrix::inverted() (adjugate / determinant)"""
    R, O = libs
    rng = np.random.default_rng(0)
    for _ in range(500):
        a, b = _rand_affine(rng), _rand_affine(rng)
        ro, oo = np.zeros(16, np.float32), np.zeros(16, np.float32)
        R.ref_mat4_mul(a.ctypes.data, b.ctypes.data, ro.ctypes.data)
        O.orc_mat4_mul(a.ctypes.data, b.ctypes.data, oo.ctypes.data)
        assert np.array_equal(ro.view(np.uint32), oo.view(np.uint32))
        R.ref_mat4_inverted(a.ctypes.data, ro.ctypes.data)
        O.orc_mat4_inverted(a.ctypes.data, oo.ctypes.data)
        assert np.array_equal(ro.view(np.uint32), oo.view(np.uint32))
        p = _f(rng.uniform(-5, 5, size=3))
        r3, o3 = np.zeros(3, np.float32), np.zeros(3, np.float32)
        R.ref_mat4_transform_point(a.ctypes.data, p.ctypes.data, r3.ctypes.data)
        O.orc_mat4_transform_point(a.ctypes.data, p.ctypes.data, o3.ctypes.data)
        assert np.array_equal(r3.view(np.uint32), o3.view(np.uint32))
        R.ref_mat4_scaling_of(a.ctypes.data, r3.ctypes.data)
        O.orc_mat4_scaling_of(a.ctypes.data, o3.ctypes.data)
        assert np.array_equal(r3.view(np.uint32), o3.view(np.uint32))
        R.ref_vec3_normalized(p.ctypes.data, r3.ctypes.data)
        O.orc_vec3_normalized(p.ctypes.data, o3.ctypes.data)
        assert np.array_equal(r3.view(np.uint32), o3.view(np.uint32))


def test_rotations_within_one_ulp(libs):
    """Matrix4::rotation / rotationX / rotationY: identical structure; entries within 1 ulp because the reference calls
    glibc's sinf/cosf while oracle and device use the correctly rounded value (DESIGN.md 'numerics')"""
    R, O = libs
    R.ref_mat4_rotation.argtypes = [C.c_float, C.c_void_p, C.c_void_p]
    O.orc_mat4_rotation.argtypes = [C.c_float, C.c_void_p, C.c_void_p]
    for f in (R.ref_mat4_rotation_x, R.ref_mat4_rotation_y, O.orc_mat4_rotation_x, O.orc_mat4_rotation_y):
        f.argtypes = [C.c_float, C.c_void_p]
    rng = np.random.default_rng(1)
    exact = total = 0
    for _ in range(2000):
        ang = float(np.float32(rng.uniform(-3.2, 3.2)))
        ax = rng.normal(size=3); ax = _f(ax / np.linalg.norm(ax))
        ro, oo = np.zeros(16, np.float32), np.zeros(16, np.float32)
        for rf, of, args in ((R.ref_mat4_rotation, O.orc_mat4_rotation, (ax.ctypes.data,)), (R.ref_mat4_rotation_x, O.orc_mat4_rotation_x, ()),
                             (R.ref_mat4_rotation_y, O.orc_mat4_rotation_y, ())):
            rf(ang, *args, ro.ctypes.data)
            of(ang, *args, oo.ctypes.data)
            assert np.array_equal(ro == 0, oo == 0)
            assert np.allclose(ro, oo, rtol=0, atol=2.5e-7), (ro, oo)
            exact += int(np.array_equal(ro.view(np.uint32), oo.view(np.uint32))); total += 1
    assert exact / total > 0.5


def test_voxel_hash_rng_and_hash_map_order(libs):
    """voxel_grid.hpp hash + toVoxel, util.hpp RNG helpers, and the iteration order of the voxel hash map that decides the
    greedy box decomposition (component_voxel_grid.hpp:114-118)"""
    R, O = libs
    R.ref_voxel_hash.restype = C.c_ulonglong
    O.orc_voxel_hash.restype = C.c_ulonglong
    rng = np.random.default_rng(2)
    for _ in range(300):
        x, y, z = (int(v) for v in rng.integers(-500, 500, size=3))
        assert R.ref_voxel_hash(x, y, z) == O.orc_voxel_hash(x, y, z)
    R.ref_to_voxel.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
    O.orc_to_voxel.argtypes = [C.c_float, C.c_float, C.c_float, C.c_void_p]
    for _ in range(300):
        p = rng.uniform(-40, 40, size=3)
        a, b = np.zeros(3, np.int32), np.zeros(3, np.int32)
        R.ref_to_voxel(*map(float, p), a.ctypes.data)
        O.orc_to_voxel(*map(float, p), b.ctypes.data)
        assert np.array_equal(a, b)
    for seed in (1, 42, 12345):
        ia, fa, ib, fb = np.zeros(64, np.int32), np.zeros(64, np.float32), np.zeros(64, np.int32), np.zeros(64, np.float32)
        R.ref_rng_stream(seed, 3, 30, 64, ia.ctypes.data, fa.ctypes.data)
        O.orc_rng_stream(seed, 3, 30, 64, ib.ctypes.data, fb.ctypes.data)
        assert np.array_equal(ia, ib) and np.array_equal(fa.view(np.uint32), fb.view(np.uint32))
    for _ in range(10):
        L, Hh, Wd = int(rng.integers(12, 30)), int(rng.integers(5, 7)), int(rng.integers(12, 25))
        pts = [(x, 0, z) for x in range(L) for z in range(Wd)] + [(0, y, z) for y in range(Hh) for z in range(Wd)] + [(x, y, 0) for x in range(L) for y in range(Hh)]
        xyz = np.array(pts, dtype=np.int32)
        oa, ob = np.zeros_like(xyz), np.zeros_like(xyz)
        ka = R.ref_voxel_grid_order(xyz.ctypes.data, len(pts), oa.ctypes.data)
        kb = O.orc_voxel_grid_order(xyz.ctypes.data, len(pts), ob.ctypes.data)
        assert ka == kb and np.array_equal(oa[:ka], ob[:kb])


@pytest.mark.parametrize("size", [2, 3, 5, 7])
def test_honeycomb_maze_matches_reference_library(libs, size):
    """the honeycomb maze + Kruskal restatement (oracle/orc_maze.hpp) against the reference's own maze library compiled from
    /root/reference, both seeded explicitly: same adjacency lists in the same order, same border segments and centres (bits)"""
    ref, L = libs
    if not hasattr(ref, "ref_honeycomb_maze"):
        pytest.skip("oracle/_ref/libmvref.so predates the maze shim")
    for fn in (ref.ref_honeycomb_maze, L.orc_honeycomb_maze):
        fn.argtypes = [C.c_int, C.c_uint, C.c_void_p, C.c_int]
        fn.restype = C.c_int
    for seed in (1, 42, 12345):
        a = np.zeros(1 << 16, dtype=np.float64); b = np.zeros(1 << 16, dtype=np.float64)
        na = ref.ref_honeycomb_maze(size, seed, a.ctypes.data, a.size)
        nb = L.orc_honeycomb_maze(size, seed, b.ctypes.data, b.size)
        assert na == nb and na > 0
        assert np.array_equal(a[:na].view(np.uint64), b[:nb].view(np.uint64))


def test_perlin_noise_matches_reference_header(libs):
    """Collect's landscape noise: the restated PerlinNoise against the reference's own header, bit for bit, over the argument range the
    scenario uses (x / fx, z / fz with frequencies 0.1 .. 9.9, 1 .. 9 octaves, seeds below 1e9)"""
    ref, L = libs
    if not hasattr(ref, "ref_perlin"):
        pytest.skip("oracle/_ref/libmvref.so predates the perlin shim")
    rng = np.random.default_rng(5)
    for fn in (ref.ref_perlin, L.orc_perlin):
        fn.argtypes = [C.c_uint, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        fn.restype = None
    for seed in (0, 1, 123456789, 999999999):
        for octaves in (1, 3, 9):
            xy = np.ascontiguousarray(rng.uniform(0, 42, size=(500, 2)) / rng.uniform(0.4, 420, size=(500, 1)))
            a = np.zeros(500); b = np.zeros(500)
            ref.ref_perlin(seed, 500, xy.ctypes.data, octaves, a.ctypes.data)
            L.orc_perlin(seed, 500, xy.ctypes.data, octaves, b.ctypes.data)
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), (seed, octaves)


def test_color_tables_match_reference(libs):
    """env/const.hpp's colour tables (all / agent / object / layout colours) and rgb(): the oracle's and the product's copies
    against the reference header, incl. the float palette the rasteriser multiplies with (bit patterns)"""
    from megaverse_b200 import capi

    ref, L = libs
    if not hasattr(ref, "ref_color_tables"):
        pytest.skip("oracle/_ref/libmvref.so predates the colour shim")
    P = capi.lib()
    outs = []
    for fn in (ref.ref_color_tables, L.orc_color_tables, P.mv_debug_color_tables):
        fn.argtypes = [C.c_void_p, C.c_int]
        fn.restype = C.c_int
        buf = np.zeros(512, dtype=np.uint32)
        n = fn(buf.ctypes.data, buf.size)
        assert n > 0
        outs.append(buf[:n].copy())
    assert np.array_equal(outs[0], outs[1]), "oracle colour tables differ from the reference's"
    assert np.array_equal(outs[0], outs[2]), "product colour tables differ from the reference's"
    # the step kernel's agent colours are palette indices into allColors
    import re
    src = open(os.path.join(ROOT, "megaverse_b200", "csrc", "step_kernel.cuh")).read()
    idx = [int(x) for x in re.search(r"agentColors\[7\] = \{([0-9, ]+)\}", src).group(1).split(",")]
    n_all = int(outs[0][0])
    all_colors, agent_colors = outs[0][4:4 + n_all], outs[0][4 + n_all:4 + n_all + 7]
    assert [int(all_colors[i]) for i in idx] == [int(c) for c in agent_colors]


def test_scene_graph_conventions_match_magnum(libs):
    """what the restatement assumes about Magnum's SceneGraph, against the real one: scale / rotateY / translate prepend, scaleLocal
    appends, a child's absolute matrix = parent * local (also through Object::setClean, the renderer's path), and
    setParentKeepTransformation = inverse(parent abs) * abs.  Bit-exact without rotation; with a rotation the sin/cos values differ
    by at most 1 ulp (the transcendentals note in DESIGN.md), so that case is compared with a tolerance."""
    ref, L = libs
    if not hasattr(ref, "ref_scenegraph_case"):
        pytest.skip("oracle/_ref/libmvref.so predates the scene-graph shim")
    fp = C.c_void_p
    for fn in (ref.ref_scenegraph_case, L.orc_scenegraph_case):
        fn.argtypes = [fp, C.c_float, fp, fp, fp, fp, fp, fp, fp, fp]
        fn.restype = None
    rng = np.random.default_rng(3)
    for trial in range(200):
        v = [_f(rng.uniform(0.1, 3.0, 3)) for _ in range(8)]
        v[1] = _f(rng.uniform(-20, 20, 3)); v[3] = _f(rng.uniform(-1, 1, 3)); v[5] = _f(rng.uniform(-20, 20, 3)); v[7] = _f(rng.uniform(-20, 20, 3))
        ps, pt, cs, ct, fs, ft, rs, rt = v
        for angle, exact in ((0.0, True), (float(rng.uniform(-3.1, 3.1)), False)):
            a = np.zeros(48, dtype=np.float32); b = np.zeros(48, dtype=np.float32)
            args = lambda out: (ps.ctypes.data, C.c_float(angle), pt.ctypes.data, cs.ctypes.data, ct.ctypes.data, fs.ctypes.data, ft.ctypes.data,
                                rs.ctypes.data, rt.ctypes.data, out.ctypes.data)
            ref.ref_scenegraph_case(*args(a)); L.orc_scenegraph_case(*args(b))
            assert np.array_equal(a[:16].view(np.uint32), a[16:32].view(np.uint32)), "setClean changes the absolute matrix"
            if exact:
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), trial
            else:
                assert np.allclose(a, b, rtol=2e-6, atol=2e-6), trial
                assert np.array_equal(a[32:].view(np.uint32), b[32:].view(np.uint32)), trial  # the re-parented object has no rotation


def test_platforms_match_reference_header(libs):
    """the obstacle-course platforms (Empty / Wall / Lava / Step / Gap / Start / Exit / Transition) of the oracle against the
    reference's own scenarios/platforms.hpp compiled in place: same RNG draws in init() / generate(), same layout, wall and terrain
    boxes after the scene-graph transforms (incl. the 90-degree turns), same object cells, spawn points, difficulty flags and anchors"""
    ref, L = libs
    if not hasattr(ref, "ref_platform_case"):
        pytest.skip("oracle/_ref/libmvref.so predates the platform shim")
    for fn in (ref.ref_platform_case, L.orc_platform_case):
        fn.argtypes = [C.c_int, C.c_uint, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        fn.restype = C.c_int
    param_sets = [_f([1, 2, 1, 4, 1, 3, 0.2, 60, 1]), _f([2, 3, 3, 10, 2, 4, 0.2, 60, 1]), _f([1, 3, 2, 10, 1, 3, 0.2, 60, 1])]
    checked = 0
    for ptype in range(8):
        for seed in range(1, 26):
            for width in (-1, 5, 8):
                if ptype == 7 and width == -1:
                    continue
                for rotate, prev_w in ((0, 0), (1, 6), (2, 7)):
                    params = param_sets[seed % 3]
                    walls = [4 | 8, 15, 1 | 4 | 8][seed % 3]
                    a = np.zeros(4096, dtype=np.int32); b = np.zeros(4096, dtype=np.int32)
                    args = (ptype, seed, walls, width, 3 + seed % 4, rotate, prev_w, params.ctypes.data, 1 + seed % 5, 1 + seed % 4)
                    na = ref.ref_platform_case(*args, a.ctypes.data, a.size)
                    nb = L.orc_platform_case(*args, b.ctypes.data, b.size)
                    assert na == nb and na > 0, (ptype, seed, width, rotate)
                    assert np.array_equal(a[:na], b[:nb]), "platform type %d seed %d width %d rotate %d: first diff at %s" % (
                        ptype, seed, width, rotate, np.nonzero(a[:na] != b[:nb])[0][:5])
                    checked += 1
    assert checked > 1500


def test_voxel_layout_pipeline_matches_reference(libs):
    """platforms -> VoxelGridComponent::addPlatform -> toBoundingBoxes (the greedy voxel merge that produces every static box the
    scenarios draw and collide with): the restatement against the reference's own component_voxel_grid.hpp + platforms.hpp, for a
    start platform followed by each obstacle type, straight and turned both ways, with visible and invisible walls.  Group order
    (std::map<BBoxInfo>), box order (hash-map iteration) and extents must be identical."""
    ref, L = libs
    if not hasattr(ref, "ref_voxel_layout_case"):
        pytest.skip("oracle/_ref/libmvref.so predates the layout shim")
    for fn in (ref.ref_voxel_layout_case, L.orc_voxel_layout_case):
        fn.argtypes = [C.c_int, C.c_uint, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        fn.restype = C.c_int
    param_sets = [_f([1, 2, 1, 4, 1, 3, 0.2, 60, 1]), _f([2, 3, 3, 10, 2, 4, 0.2, 60, 1])]
    checked = 0
    for ptype in (1, 2, 3, 4, 6):
        for seed in range(1, 41):
            for rotate in (0, 1, 2):
                a = np.zeros(1 << 14, dtype=np.int32); b = np.zeros(1 << 14, dtype=np.int32)
                args = (ptype, seed, rotate, seed % 2, param_sets[seed % 2].ctypes.data)
                na = ref.ref_voxel_layout_case(*args, a.ctypes.data, a.size)
                nb = L.orc_voxel_layout_case(*args, b.ctypes.data, b.size)
                assert na == nb and na > 4, (ptype, seed, rotate)
                assert np.array_equal(a[:na], b[:nb]), "type %d seed %d rotate %d: first diff at %s" % (ptype, seed, rotate, np.nonzero(a[:na] != b[:nb])[0][:5])
                checked += 1
    assert checked == 600


def test_object_stacking_matches_reference_header(libs):
    """ObjectStackingComponent::onInteractAction (pick up from the voxel in front / the one above it unless something sits on top,
    carry as a child of the pickup spot at 0.78 scale, put down into the voxel of the carried object's position, let it sink to the
    first solid / occupied voxel, refuse occupied voxels and voxels holding another agent): the oracle's Env::onInteractAction
    against the reference's own component_object_stacking.hpp driven on real Magnum scene-graph objects.  Random scripts of agent
    poses; object poses, parents, collision flags, carrying state and voxel occupancy must agree bit for bit after every event."""
    ref, L = libs
    if not hasattr(ref, "ref_stacking_case"):
        pytest.skip("oracle/_ref/libmvref.so predates the stacking shim")
    vp = C.c_void_p
    for fn in (ref.ref_stacking_case, L.orc_stacking_case):
        fn.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, vp, C.c_int]
        fn.restype = C.c_int

    def rot_y(a):
        c, s = np.float32(np.cos(a)), np.float32(np.sin(a))
        m = np.eye(4, dtype=np.float32); m[0, 0] = c; m[0, 2] = -s; m[2, 0] = s; m[2, 2] = c  # column-major rows = columns
        return m

    def rot_x(a):
        c, s = np.float32(np.cos(a)), np.float32(np.sin(a))
        m = np.eye(4, dtype=np.float32); m[1, 1] = c; m[1, 2] = s; m[2, 1] = -s; m[2, 2] = c
        return m

    def trans(t):
        m = np.eye(4, dtype=np.float32); m[3, :3] = t
        return m

    rng = np.random.default_rng(17)
    total_picks = 0
    for trial in range(60):
        solid = np.array([[x, 0, z] for x in range(8) for z in range(8)] + [[4, 1, 4], [4, 2, 4], [1, 1, 6]], dtype=np.int32)
        cells = [(2, 1, 2), (2, 2, 2), (5, 1, 3), (6, 1, 6), (3, 1, 5), (5, 1, 5)]
        objs = np.array(cells[: 3 + trial % 4], dtype=np.int32)
        A = 1 + trial % 3
        events = []
        for ev in range(40):
            ai = int(rng.integers(0, A))
            if rng.random() < 0.6:  # stand next to an object / target cell and face it
                tx, ty, tz = cells[int(rng.integers(0, len(cells)))]
                yaw = float(rng.uniform(-np.pi, np.pi))
                fwd = np.array([-np.sin(yaw), 0.0, -np.cos(yaw)])  # the pickup spot is 1 m along the agent's -z
                pos = np.array([tx + 0.5, 1.0 + 1.8 + rng.uniform(-0.1, 0.6), tz + 0.5]) - fwd * rng.uniform(0.7, 1.3)
            else:
                yaw = float(rng.uniform(-np.pi, np.pi))
                pos = np.array([rng.uniform(0.5, 7.5), 1.0 + 1.8 + rng.uniform(-0.2, 1.5), rng.uniform(0.5, 7.5)])
            agent_m = rot_y(yaw) @ trans(pos.astype(np.float32))          # row-vector convention of the flattened column-major matrix
            cam_m = rot_x(float(rng.uniform(-0.2, 0.2))) @ trans(np.float32([0, 0.41, 0]))
            events.append(np.concatenate([[np.float32(ai)], agent_m.ravel(), cam_m.ravel()]))
        script = np.ascontiguousarray(np.array(events, dtype=np.float32))
        a = np.zeros(1 << 16, dtype=np.int32); b = np.zeros(1 << 16, dtype=np.int32)
        args = (solid.ctypes.data, len(solid), objs.ctypes.data, len(objs), A, script.ctypes.data, len(events))
        na = ref.ref_stacking_case(*args, a.ctypes.data, a.size)
        nb = L.orc_stacking_case(*args, b.ctypes.data, b.size)
        assert na == nb and na > 0, trial
        assert np.array_equal(a[:na], b[:nb]), "trial %d: first diff at %s" % (trial, np.nonzero(a[:na] != b[:nb])[0][:5])
        pos_ = 0
        for _ in range(len(events)):  # rows: objects * 8, carrying per agent, count, count * 4
            carried = a[pos_ + len(objs) * 8: pos_ + len(objs) * 8 + A]
            total_picks += int((carried >= 0).any())
            cnt = int(a[pos_ + len(objs) * 8 + A])
            pos_ += len(objs) * 8 + A + 1 + 4 * cnt
        assert pos_ == na
    assert total_picks > 100  # the scripts really pick things up


def test_layout_utils_match_reference(libs):
    """the drawables and colliders that layout_utils.cpp creates -- addBoundingBoxes (voxel sizes 1 and 2, drawn / solid / both),
    addTerrain, addStaticCollidingBox, addDiamond, addPillar (caps re-parented keeping their transformation), addSphere -- from the
    reference's own source file compiled in place, against the oracle: model matrices (absoluteTransformationMatrix), colours and the
    collider origin / scaling RigidBody::syncPose would hand to Bullet, bit for bit, in creation order per mesh type"""
    ref, L = libs
    if not hasattr(ref, "ref_layout_utils_case"):
        pytest.skip("oracle/_ref/libmvref.so predates the layout-utils shim")
    for fn in (ref.ref_layout_utils_case, L.orc_layout_utils_case):
        fn.argtypes = [C.c_uint, C.c_void_p, C.c_int]
        fn.restype = C.c_int
    for seed in range(1, 201):
        a = np.zeros(1 << 14, dtype=np.int32); b = np.zeros(1 << 14, dtype=np.int32)
        na = ref.ref_layout_utils_case(seed, a.ctypes.data, a.size)
        nb = L.orc_layout_utils_case(seed, b.ctypes.data, b.size)
        assert na == nb and na > 100, seed
        assert np.array_equal(a[:na], b[:nb]), "seed %d: first diff at %s (counts %s vs %s)" % (seed, np.nonzero(a[:na] != b[:nb])[0][:5], a[:5], b[:5])


# ---------------------------------------------------------------------------------------------------------------------
# The reference's REAL scenario sources, driven tick by tick beside the oracle (oracle/_ref/libmvscen.so: scenario_*.cpp,
# component_*.hpp/.cpp, layout_utils.cpp, env/scenario.hpp and scenario_default.hpp compiled in place against Bullet-free
# stand-ins, oracle/ref_shim/scen_shim.cpp).  The agents on the reference side are posed puppets that receive the oracle's
# agent transforms every tick, so what is compared -- bit for bit -- is everything a scenario decides: the level it builds at
# reset (every drawable's mesh, colour and absolute matrix in draw order, every collider handed to Bullet, the agents' spawn
# arguments, the episode length), and per tick the rewards, timers, done flag, true objectives, teleports and the moved
# drawables / toggled colliders.
SCEN = os.path.join(ROOT, "oracle", "_ref", "libmvscen.so")
MAZE_SEED_XOR = 0x6D617A65  # the oracle seeds the maze's Kruskal generator from the episode seed (upstream: std::random_device)


@pytest.fixture(scope="module")
def scen_libs(built):
    import orc

    if not os.path.exists(SCEN):
        if os.path.isdir("/root/reference/src/libs/scenarios"):
            import subprocess

            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "scen"])
        else:
            pytest.skip("oracle/_ref/libmvscen.so not built and /root/reference absent")
    R, O = C.CDLL(SCEN), orc.lib()
    R.ref_scen_create.restype = C.c_void_p
    R.ref_scen_create.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
    R.ref_scen_destroy.argtypes = [C.c_void_p]
    R.ref_scen_seed.argtypes = [C.c_void_p, C.c_int]
    R.ref_scen_reset_begin.argtypes = [C.c_void_p, C.c_uint, C.c_void_p]
    R.ref_scen_reset_end.argtypes = [C.c_void_p, C.c_void_p]
    R.ref_scen_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    R.ref_scen_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    O.orc_scen_reset.argtypes = [C.c_void_p, C.c_int]
    O.orc_scen_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    O.orc_scen_spawns.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    O.orc_scen_warp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
    O.orc_scen_poses.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    O.orc_scen_teleports.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    O.orc_scenario_dump.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return R, O


_SCEN_CAP = 1 << 20


def _scen_dump(fn, *args):
    buf = np.zeros(_SCEN_CAP, np.uint32)
    n = fn(*args, buf.ctypes.data, _SCEN_CAP)
    assert n > 0
    d = buf[:n]
    A = int(d[3])
    out = {"length": d[0], "done": int(d[1]), "sec": d[2], "agents": d[4:4 + 3 * A].copy()}
    i = 4 + 3 * A
    n_inst = int(d[i]); i += 1
    out["inst"] = d[i:i + 18 * n_inst].reshape(n_inst, 18).copy(); i += 18 * n_inst
    n_col = int(d[i]); i += 1
    out["col"] = d[i:i + 9 * n_col].reshape(n_col, 9).copy(); i += 9 * n_col
    assert i == n
    return out


def _scen_same(r, o, where):
    assert r["length"] == o["length"], f"{where}: episode length"
    assert r["done"] == o["done"], f"{where}: done"
    assert r["sec"] == o["sec"], f"{where}: episode clock {r['sec']:#x} vs {o['sec']:#x}"
    assert np.array_equal(r["agents"], o["agents"]), f"{where}: rewards / objectives {r['agents'].view(np.float32)} vs {o['agents'].view(np.float32)}"
    assert r["col"].shape == o["col"].shape, f"{where}: collider count"
    if len(r["col"]):
        # an axis-aligned box reaches Bullet with Matrix4::rotation() of its scene-graph matrix, whose diagonal is s * (1 / s): that
        # is 1 or 1 - 1ulp.  The oracle treats such boxes as exactly axis aligned; rotated ones (maze walls) must agree bit for bit.
        aligned = (o["col"][:, 6] == np.float32(1).view(np.uint32)) & (o["col"][:, 7] == 0)
        rx = r["col"][:, 6:8].copy().view(np.float32)
        assert np.all(np.abs(rx[aligned] - np.float32([1, 0])) <= 6e-8), f"{where}: an axis-aligned collider is not"
        r = dict(r); r["col"] = r["col"].copy(); r["col"][aligned, 6:8] = o["col"][aligned, 6:8]
    for k in ("inst", "col"):
        assert r[k].shape == o[k].shape, f"{where}: {k} count {r[k].shape[0]} vs {o[k].shape[0]}"
        if not np.array_equal(r[k], o[k]):
            bad = np.nonzero((r[k] != o[k]).any(axis=1))[0]
            raise AssertionError(f"{where}: {k} rows {bad[:8]} of {len(r[k])} differ, first ref {r[k][bad[0]]} oracle {o[k][bad[0]]}")


def _cosim(R, O, scenario, A, seed, max_ticks, episodes=2, params=None, warp_every=0):
    import helpers
    import orc

    params = params or {}
    keys = (C.c_char_p * max(1, len(params)))(*[k.encode() for k in params])
    vals = (C.c_float * max(1, len(params)))(*[float(v) for v in params.values()])
    o = orc.Oracle(scenario, 1, A, params=params, render=False)
    rh = R.ref_scen_create(scenario.encode(), A, keys, vals, len(params))
    assert rh
    stats = {"ticks": 0, "reward_events": 0, "dones": 0, "teleports": 0}
    try:
        o.seed_env(0, seed)
        R.ref_scen_seed(rh, seed)
        rng = np.random.default_rng(seed)
        poses = np.zeros(33 * A, np.float32)
        for ep in range(episodes):
            O.orc_scen_reset(o.h_, 0)
            sp_r, sp_o = np.zeros(4 * A, np.uint32), np.zeros(4 * A, np.uint32)
            R.ref_scen_reset_begin(rh, MAZE_SEED_XOR, sp_r.ctypes.data)
            O.orc_scen_spawns(o.h_, 0, sp_o.ctypes.data)
            assert np.array_equal(sp_r, sp_o), f"{scenario} ep {ep}: spawn arguments {sp_r.view(np.float32)} vs {sp_o.view(np.float32)}"
            O.orc_scen_poses(o.h_, 0, poses.ctypes.data)
            R.ref_scen_reset_end(rh, poses.ctypes.data)
            _scen_same(_scen_dump(R.ref_scen_dump, rh), _scen_dump(O.orc_scenario_dump, o.h_, 0), f"{scenario} A={A} seed={seed} ep={ep} reset")
            last = None
            for t in range(max_ticks):
                if warp_every and last is not None and t % warp_every == warp_every - 1:
                    # drop agents beside random drawables (objects, boxes, rewards, walls ...) so that the scripted walk interacts
                    # with the level far more often than it would on foot
                    world = last["inst"][last["inst"][:, 14].view(np.float32) < 400]
                    special = world[(world[:, 0] != 0) | (world[:, 1] == 0x3A7FA6)]  # rewards, objects, pillars; Sokoban's boxes
                    for a in range(A):
                        if rng.random() < 0.7:
                            pool = special if len(special) and rng.random() < 0.6 else world
                            m = pool[rng.integers(len(pool)), 2:].view(np.float32)
                            if scenario == "Sokoban":  # the cell next to a box, facing one of the four directions
                                k = int(rng.integers(4))
                                dx, dz = [(1.5, 0), (-1.5, 0), (0, 1.5), (0, -1.5)][k]
                                yaw = float(rng.integers(4)) * np.pi / 2
                                O.orc_scen_warp(o.h_, 0, a, float(m[12] + dx), float(m[13] + 1.4), float(m[14] + dz), yaw)
                            else:
                                O.orc_scen_warp(o.h_, 0, a, float(m[12] + rng.uniform(-0.8, 0.8)), float(m[13] + 1.2), float(m[14] + rng.uniform(-0.8, 0.8)),
                                                float(rng.uniform(0, 2 * np.pi)))
                if scenario == "Rearrange" and ep == 0 and not warp_every:
                    acts = helpers.rearrange_controller(o, 0, A)
                else:
                    acts = np.asarray(helpers.purposeful_actions(rng, A, t), np.int32)
                acts = np.ascontiguousarray(acts, np.int32)
                if scenario == "Sokoban" and warp_every:
                    acts |= 1 << 8  # keep pushing
                O.orc_scen_step(o.h_, 0, acts.ctypes.data)
                O.orc_scen_poses(o.h_, 0, poses.ctypes.data)
                tp_r, tp_o = np.zeros(4 * A, np.uint32), np.zeros(4 * A, np.uint32)
                R.ref_scen_step(rh, acts.ctypes.data, poses.ctypes.data, tp_r.ctypes.data)
                O.orc_scen_teleports(o.h_, 0, tp_o.ctypes.data)
                assert np.array_equal(tp_r, tp_o), f"{scenario} ep {ep} t {t}: teleports {tp_r} vs {tp_o}"
                r, d = _scen_dump(R.ref_scen_dump, rh), _scen_dump(O.orc_scenario_dump, o.h_, 0)
                _scen_same(r, d, f"{scenario} A={A} seed={seed} ep={ep} t={t}")
                stats["ticks"] += 1
                stats["reward_events"] += int((r["agents"].view(np.float32)[0::3] != 0).sum())
                stats["teleports"] += int(tp_r[0::4].sum())
                last = r
                if r["done"]:
                    stats["dones"] += 1
                    break
    finally:
        o.close()
        R.ref_scen_destroy(rh)
    return stats


@pytest.mark.parametrize("scenario,A,seed,ticks", [
    ("TowerBuilding", 1, 3, 1400), ("TowerBuilding", 4, 17, 500),
    ("ObstaclesEasy", 2, 5, 600), ("ObstaclesMedium", 1, 9, 600), ("ObstaclesHard", 3, 13, 600),
    ("ObstaclesWalls", 2, 21, 300), ("ObstaclesSteps", 2, 22, 300), ("ObstaclesLava", 2, 23, 300),
    ("Collect", 1, 4, 1200), ("Collect", 4, 8, 500),
    ("Sokoban", 1, 6, 700), ("Sokoban", 2, 12, 400),
    ("Rearrange", 1, 10, 900), ("Rearrange", 3, 11, 400),
    ("HexExplore", 1, 14, 700), ("HexExplore", 4, 15, 300),
    ("HexMemory", 1, 16, 900), ("HexMemory", 2, 18, 400),
])
def test_real_scenario_sources_match_the_oracle_tick_by_tick(scen_libs, scenario, A, seed, ticks):
    R, O = scen_libs
    stats = _cosim(R, O, scenario, A, seed, ticks)
    assert stats["ticks"] > 0


@pytest.mark.parametrize("scenario,A,seed,ticks,warp", [
    ("TowerBuilding", 1, 3, 1400, 25), ("TowerBuilding", 4, 17, 500, 25),
    ("ObstaclesEasy", 2, 5, 600, 25), ("ObstaclesHard", 3, 13, 600, 25),
    ("Collect", 4, 8, 500, 25),
    ("Sokoban", 2, 9, 700, 12), ("Sokoban", 1, 4, 700, 12),
    ("Rearrange", 2, 10, 600, 25),
    ("HexExplore", 1, 14, 700, 25),
    ("HexMemory", 1, 16, 900, 25), ("HexMemory", 2, 18, 400, 25),
])
def test_real_scenario_sources_match_the_oracle_with_agents_dropped_next_to_things(scen_libs, scenario, A, seed, ticks, warp):
    """same comparison on 40-second episodes (the timer runs out inside the window) with agents put beside random objects, boxes and
    rewards every `warp` ticks: pick-ups, placements, pushes, collections and both ways of finishing an episode all occur"""
    R, O = scen_libs
    stats = _cosim(R, O, scenario, A, seed, ticks, warp_every=warp, params={"episodeLengthSec": 40.0})
    assert stats["reward_events"] + stats["dones"] > 0, stats


# ---------------------------------------------------------------------------------------------------------------------
# The reference's WHOLE env library running on its own (oracle/_ref/libmvenv.so): env.cpp, agent.cpp, the character controller,
# RigidBody / MotionState and every scenario source, compiled unmodified against a stand-in for the absent Bullet
# (oracle/ref_shim/mini_bullet: LinearMath and the world's containers restated, narrow phase = the oracle's analytic definitions).
# No puppets here: the reference's own kinematics drive its agents; the oracle has to follow bit for bit, tick after tick.
ENVLIB = os.path.join(ROOT, "oracle", "_ref", "libmvenv.so")


@pytest.fixture(scope="module")
def env_libs(built):
    import orc

    if not os.path.exists(ENVLIB):
        if os.path.isdir("/root/reference/src/libs/env"):
            import subprocess

            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "envlib"])
        else:
            pytest.skip("oracle/_ref/libmvenv.so not built and /root/reference absent")
    R, O = C.CDLL(ENVLIB), orc.lib()
    R.ref_env_create.restype = C.c_void_p
    R.ref_env_create.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
    R.ref_env_destroy.argtypes = [C.c_void_p]
    R.ref_env_seed.argtypes = [C.c_void_p, C.c_int]
    R.ref_env_reset.argtypes = [C.c_void_p, C.c_uint]
    R.ref_env_step.argtypes = [C.c_void_p, C.c_void_p]
    R.ref_env_warp.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
    R.ref_env_dump.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    R.ref_env_views.argtypes = [C.c_void_p, C.c_void_p]
    O.orc_scen_reset.argtypes = [C.c_void_p, C.c_int]
    O.orc_scen_step.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    O.orc_scen_warp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float]
    O.orc_scenario_dump.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    return R, O


def _full_run(R, O, scenario, A, seed, max_ticks, episodes=2, params=None, warp_every=0):
    import helpers
    import orc

    params = params or {}
    keys = (C.c_char_p * max(1, len(params)))(*[k.encode() for k in params])
    vals = (C.c_float * max(1, len(params)))(*[float(v) for v in params.values()])
    o = orc.Oracle(scenario, 1, A, params=params, render=False)
    rh = R.ref_env_create(scenario.encode(), A, keys, vals, len(params))
    assert rh
    stats = {"ticks": 0, "reward_events": 0, "dones": 0}
    try:
        o.seed_env(0, seed)
        R.ref_env_seed(rh, seed)
        rng = np.random.default_rng(seed)
        for ep in range(episodes):
            O.orc_scen_reset(o.h_, 0)
            R.ref_env_reset(rh, MAZE_SEED_XOR)
            if O.orc_scen_undefined_spawn(o.h_, 0):
                # fewer spawn points than agents on a small start platform: the reference indexes past the end of its vector
                # (platforms.hpp:221-244, scenario_default.hpp:83-91) and the agent lands wherever the heap says -- nothing to compare
                stats["undefined_spawn"] = stats.get("undefined_spawn", 0) + 1
                break
            last = _scen_dump(R.ref_env_dump, rh)
            _scen_same(last, _scen_dump(O.orc_scenario_dump, o.h_, 0), f"{scenario} A={A} seed={seed} ep={ep} reset")
            for t in range(max_ticks):
                if warp_every and t % warp_every == warp_every - 1:
                    world = last["inst"][last["inst"][:, 14].view(np.float32) < 400]
                    special = world[(world[:, 0] != 0) | (world[:, 1] == 0x3A7FA6)]
                    for a in range(A):
                        if rng.random() < 0.7:
                            pool = special if len(special) and rng.random() < 0.6 else world
                            m = pool[rng.integers(len(pool)), 2:].view(np.float32)
                            if scenario == "Sokoban":
                                dx, dz = [(1.5, 0), (-1.5, 0), (0, 1.5), (0, -1.5)][int(rng.integers(4))]
                                pos = (float(m[12] + dx), float(m[13] + 1.4), float(m[14] + dz), float(rng.integers(4)) * np.pi / 2)
                            else:
                                pos = (float(m[12] + rng.uniform(-0.8, 0.8)), float(m[13] + 1.2), float(m[14] + rng.uniform(-0.8, 0.8)), float(rng.uniform(0, 2 * np.pi)))
                            O.orc_scen_warp(o.h_, 0, a, *pos)
                            R.ref_env_warp(rh, a, *pos)
                if scenario == "Rearrange" and ep == 0 and not warp_every:
                    acts = helpers.rearrange_controller(o, 0, A)
                else:
                    acts = np.asarray(helpers.purposeful_actions(rng, A, t), np.int32)
                acts = np.ascontiguousarray(acts, np.int32)
                if scenario == "Sokoban" and warp_every:
                    acts |= 1 << 8
                O.orc_scen_step(o.h_, 0, acts.ctypes.data)
                R.ref_env_step(rh, acts.ctypes.data)
                last = _scen_dump(R.ref_env_dump, rh)
                _scen_same(last, _scen_dump(O.orc_scenario_dump, o.h_, 0), f"{scenario} A={A} seed={seed} ep={ep} t={t}")
                views = np.zeros((A, 16), np.float32)
                R.ref_env_views(rh, views.ctypes.data)  # Camera3D::cameraMatrix(): what the renderer takes as the view matrix
                for a in range(A):
                    assert np.array_equal(views[a].view(np.uint32), o.view(0, a).view(np.uint32)), f"{scenario} ep={ep} t={t}: view matrix of agent {a}"
                stats["ticks"] += 1
                stats["reward_events"] += int((last["agents"].view(np.float32)[0::3] != 0).sum())
                if last["done"]:
                    stats["dones"] += 1
                    break
    finally:
        o.close()
        R.ref_env_destroy(rh)
    return stats


@pytest.mark.parametrize("scenario,A,seed,ticks,warp,params", [
    ("TowerBuilding", 1, 3, 1400, 0, None), ("TowerBuilding", 4, 17, 600, 0, None), ("TowerBuilding", 2, 31, 700, 25, {"episodeLengthSec": 40.0}),
    ("ObstaclesEasy", 2, 5, 600, 0, None), ("ObstaclesMedium", 1, 9, 800, 0, None), ("ObstaclesHard", 3, 13, 600, 0, None),
    ("ObstaclesHard", 2, 14, 700, 25, {"episodeLengthSec": 40.0}),
    ("ObstaclesWalls", 2, 21, 300, 0, None), ("ObstaclesSteps", 2, 22, 300, 0, None), ("ObstaclesLava", 2, 23, 500, 0, None),
    ("Collect", 1, 4, 1200, 0, None), ("Collect", 4, 8, 600, 25, {"episodeLengthSec": 40.0}),
    ("Sokoban", 1, 6, 700, 0, None), ("Sokoban", 2, 9, 700, 12, {"episodeLengthSec": 40.0}),
    ("Rearrange", 1, 10, 900, 0, None), ("Rearrange", 2, 10, 600, 25, {"episodeLengthSec": 40.0}),
    ("HexExplore", 1, 14, 700, 0, None), ("HexExplore", 2, 19, 700, 25, {"episodeLengthSec": 40.0}),
    ("HexMemory", 1, 16, 900, 0, None), ("HexMemory", 2, 18, 500, 25, {"episodeLengthSec": 40.0}),
    # non-default parameters: a wide camera pitch range, a custom obstacle course, eight agents
    ("TowerBuilding", 2, 41, 400, 0, {"verticalLookLimitRad": 0.9}),
    ("ObstaclesEasy", 2, 42, 500, 25, {"obstaclesMinNumPlatforms": 3, "obstaclesMaxNumPlatforms": 5, "obstaclesMinGap": 2, "obstaclesMaxGap": 4, "obstaclesMinLava": 2,
                                       "obstaclesMaxLava": 6, "obstaclesMinHeight": 1, "obstaclesMaxHeight": 4, "obstaclesNumAllowedMaxDifficulty": 2}),
    ("Collect", 8, 43, 300, 25, {"episodeLengthSec": 30.0}),
    ("TowerBuilding", 8, 44, 300, 25, None),
    # the debugging scenario of the reference README's performance figure: one box, every agent spawned at the same spot
    ("Empty", 1, 45, 400, 0, {"episodeLengthSec": 10.0}), ("Empty", 3, 46, 400, 25, {"episodeLengthSec": 10.0}),
])
def test_reference_env_library_on_stand_in_bullet_matches_the_oracle(env_libs, scenario, A, seed, ticks, warp, params):
    R, O = env_libs
    stats = _full_run(R, O, scenario, A, seed, ticks, params=params, warp_every=warp)
    assert stats["ticks"] > 0


@pytest.mark.parametrize("scenario", ["TowerBuilding", "ObstaclesEasy", "ObstaclesMedium", "ObstaclesHard", "ObstaclesWalls", "ObstaclesSteps", "ObstaclesLava",
                                      "Collect", "Sokoban", "Rearrange", "HexExplore", "HexMemory", "Empty"])
def test_default_reward_shaping_and_parameters_match_the_reference_scenarios(env_libs, scenario):
    """Scenario::init() of the real scenario classes (reward shaping incl. teamSpirit, float parameters) against the tables the product
    builds its engine from (host-only accessor, no GPU needed)"""
    from megaverse_b200 import capi

    R, _ = env_libs
    R.ref_env_defaults.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    rh = R.ref_env_create(scenario.encode(), 2, None, None, 0)
    buf = C.create_string_buffer(8192)
    assert R.ref_env_defaults(rh, buf, 8192) > 0
    R.ref_env_destroy(rh)
    ref = buf.value.decode()
    L = capi.lib()
    L.mv_debug_defaults.argtypes = [C.c_char_p, C.c_char_p, C.c_int]
    buf2 = C.create_string_buffer(8192)
    assert L.mv_debug_defaults(scenario.encode(), buf2, 8192) > 0
    ours = buf2.value.decode()
    ref_r = sorted(l for l in ref.splitlines() if l.startswith("R "))
    our_r = sorted(l for l in ours.splitlines() if l.startswith("R "))
    assert ref_r == our_r, "reward shaping:\nreference %s\nproduct   %s" % (ref_r, our_r)
    # float parameters: every parameter the reference scenario defines must exist with the same default
    ref_p = dict(l[2:].split("=") for l in ref.splitlines() if l.startswith("P "))
    our_p = dict(l[2:].split("=") for l in ours.splitlines() if l.startswith("P "))
    assert ref_p == {k: our_p.get(k) for k in ref_p}, "parameters:\nreference %s\nproduct   %s" % (ref_p, our_p)


# ---------------------------------------------------------------------------------------------------------------------
# The renderer's shading, from the shader's own text: oracle/_ref/libmvshade.so compiles `compute_color()` cut out of V4R's uber.frag
# (at build time, see oracle/Makefile) with the glm vendored next to it, and the vertex stage's normal matrix with the same glm calls.
SHADE = os.path.join(ROOT, "oracle", "_ref", "libmvshade.so")


@pytest.fixture(scope="module")
def shade_libs(built):
    import orc

    if not os.path.exists(SHADE):
        if os.path.isdir("/root/reference/src/3rdparty/v4r"):
            import subprocess

            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "shade"])
        else:
            pytest.skip("oracle/_ref/libmvshade.so not built and /root/reference absent")
    return C.CDLL(SHADE), orc.lib()


def test_fragment_stage_matches_the_shader_text(shade_libs):
    """uber.frag:112-141 evaluated by glm on the CPU vs the oracle's restatement: camera-space positions all over the frustum, normals of
    any length and direction (incl. facing away and grazing), every palette colour.  The oracle computes pow(x, 300) by repeated squaring
    and GLSL leaves pow's precision to the implementation, so the colours agree to a few float ulps of the sum, and the bytes written to
    the R8G8B8A8_UNORM attachment to +-1 -- the bar BASELINE.json sets for RGB."""
    R, O = shade_libs
    rng = np.random.default_rng(3)
    n = 200000
    P = np.stack([rng.uniform(-40, 40, n), rng.uniform(-25, 25, n), -np.exp(rng.uniform(np.log(0.02), np.log(110), n))], 1).astype(np.float32)
    N = (rng.normal(size=(n, 3)) * np.exp(rng.uniform(-3, 3, (n, 1)))).astype(np.float32)
    # a share of highlights: normals close to the half-way direction between light and eye
    k = n // 4
    cd = -P[:k].astype(np.float64)
    ld = np.array([0.0, 4.0, 2.0]) + cd
    h = ld / np.linalg.norm(ld, axis=1, keepdims=True) + cd / np.linalg.norm(cd, axis=1, keepdims=True)
    N[:k] = (h / np.linalg.norm(h, axis=1, keepdims=True) + rng.normal(scale=0.02, size=(k, 3))).astype(np.float32)
    pal = np.array([[(c >> 16) & 255, (c >> 8) & 255, c & 255] for c in
                    [0xffdd3c, 0x3bb372, 0x50c878, 0x2eb5d0, 0xadd8e6, 0x3a7fa6, 0x2c3e50, 0xffb400, 0xb3b3b3, 0x555555, 0x222222, 0xffffff, 0xff0000, 0xffa770,
                     0xd468ee, 0xffe6e6, 0xffffe6, 0xccffcc, 0xe6ecff, 0xd9d9d9, 0xf2e6ff, 0xffebcc]], np.float32) / np.float32(255.0)
    D = np.ascontiguousarray(pal[rng.integers(len(pal), size=n)])
    P, N = np.ascontiguousarray(P), np.ascontiguousarray(N)
    ref, ours, ours8 = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 3), np.uint8)
    R.ref_shade(n, C.c_void_p(P.ctypes.data), C.c_void_p(N.ctypes.data), C.c_void_p(D.ctypes.data), C.c_void_p(ref.ctypes.data))
    O.orc_shade(n, C.c_void_p(P.ctypes.data), C.c_void_p(N.ctypes.data), C.c_void_p(D.ctypes.data), C.c_void_p(ours.ctypes.data), C.c_void_p(ours8.ctypes.data))
    assert np.isfinite(ref).all() and np.isfinite(ours).all()
    lit = (ref[:, 0] > 0.33 * D[:, 0] + 1e-6).mean()
    shiny = ((ref - (0.33 * D + 0.73 * D * 0.66)).max(axis=1) > 0.01).mean()  # fragments with a visible specular term
    assert lit > 0.3 and shiny > 0.02, (lit, shiny)
    assert np.abs(ref - ours).max() < 2e-4, "colour difference %g" % np.abs(ref - ours).max()
    ref8 = np.floor(np.clip(ref, 0, 1) * np.float32(255.0) + np.float32(0.5)).astype(np.int16)  # UNORM8 conversion of the attachment
    d8 = np.abs(ref8 - ours8.astype(np.int16))
    assert d8.max() <= 1, "byte difference %d" % d8.max()
    assert (d8 == 0).mean() > 0.999, "share of identical bytes %.5f" % (d8 == 0).mean()


def test_normal_matrix_matches_the_vertex_stage(shade_libs):
    """uber.vert:86: transpose(inverse(mat3(mv))) by glm vs the oracle's normalMatrix (cofactors / determinant), on model-view matrices
    of the kind the scene graph produces (rotation * non-uniform scale)"""
    R, O = shade_libs
    rng = np.random.default_rng(4)
    worst = 0.0
    for _ in range(3000):
        mv = _rand_affine(rng)
        a, b = np.zeros(9, np.float32), np.zeros(9, np.float32)
        R.ref_normal_matrix(C.c_void_p(mv.ctypes.data), C.c_void_p(a.ctypes.data))
        O.orc_normal_matrix(C.c_void_p(mv.ctypes.data), C.c_void_p(b.ctypes.data))
        scale = np.abs(a).max()
        worst = max(worst, float(np.abs(a - b).max() / scale))
    assert worst < 2e-6, worst  # same formula, different association of the 3x3 cofactor products: a few ulps


def test_action_encoding_follows_the_reference_action_space(env_libs):
    """Env::actionSpaceSizes as compiled from env.cpp, and the Action bits Env::step tests (env.hpp:22-42, through the real enum): head h
    with choice a > 0 sets bit 1 + sum_{j<h}(size_j - 1) + (a - 1), the rule of MegaverseGym::setActions (megaverse.cpp:100-116) --
    the product's mv_encode_action must agree for every combination"""
    import itertools

    from megaverse_b200 import capi

    R, _ = env_libs
    sizes = (C.c_int * 16)()
    n = R.ref_action_space_sizes(sizes, 16)
    sizes = [sizes[i] for i in range(n)]
    assert sizes == [3, 3, 3, 2, 2, 3]
    L = capi.lib()
    L.mv_encode_action.argtypes = [C.c_void_p]
    L.mv_encode_action.restype = C.c_int32
    for heads in itertools.product(*[range(s) for s in sizes]):
        want, base = 0, 0
        for h, a in enumerate(heads):
            if a > 0:
                want |= 1 << (base + a)
            base += sizes[h] - 1
        arr = np.array(heads, np.int32)
        assert L.mv_encode_action(arr.ctypes.data) == want, heads
    assert base == 10  # bits 1..10: Left, Right, Forward, Backward, LookLeft, LookRight, Jump, Interact, LookDown, LookUp
