// ORACLE tooling (test infrastructure): C entry points into pieces of the REAL reference compiled in place from
// /root/reference by oracle/Makefile (`make ref` -> oracle/_ref/libmvref.so): vendored Magnum math / primitives and the
// reference's own util headers (voxel hash, RNG helpers).  Used by tests/test_ref_shim.py to pin the restatement in
// oracle/orc_math.hpp, oracle/orc_level.hpp and oracle/orc_maze.hpp.  Nothing here is copied: the reference sources are
// included / compiled where they lie.
#include <algorithm>
#include <array>
#include <functional>
#include <cstring>
#include <vector>

#include <env/const.hpp>           // src/libs/env/include/env/const.hpp: ColorRgb, colour tables, rgb()
#include <scenarios/platforms.hpp>  // src/libs/scenarios/include/scenarios/platforms.hpp (with ref_shim/inc/env/env.hpp standing in for env.hpp)
#include <scenarios/layout_utils.hpp>  // src/libs/scenarios/src/layout_utils.cpp is compiled into this library as is
#include <scenarios/component_object_stacking.hpp>  // the pick-up / put-down logic (with the stand-ins in ref_shim/inc)
#include <scenarios/component_voxel_grid.hpp>  // VoxelGridComponent::addPlatform + the greedy voxel -> box merge (with the stand-ins in ref_shim/inc)
#include <util/perlin_noise.hpp>   // src/libs/util/include/util/perlin_noise.hpp (siv::PerlinNoise as vendored by the reference)
#include <mazes/honeycombmaze.h>  // src/libs/mazes
#include <mazes/kruskal.h>
#include <random>

#include <Magnum/Magnum.h>
#include <Magnum/Math/Matrix4.h>
#include <Magnum/Math/Vector3.h>
#include <Magnum/Primitives/Cube.h>
#include <Magnum/Trade/MeshData.h>
#include <Magnum/SceneGraph/MatrixTransformation3D.h>
#include <Magnum/SceneGraph/Object.h>
#include <Magnum/SceneGraph/Scene.h>

#include <util/util.hpp>        // src/libs/util/include/util/util.hpp: Rng, randRange, frand
#include <util/voxel_grid.hpp>  // src/libs/util/include/util/voxel_grid.hpp: VoxelCoords hash, toVoxel, VoxelGrid

using namespace Magnum;

static Matrix4 load(const float *p) { return Matrix4::from(p); }
static void store(const Matrix4 &m, float *p) { std::memcpy(p, m.data(), 64); }

extern "C" {

void ref_mat4_mul(const float *a, const float *b, float *out) { store(load(a) * load(b), out); }
void ref_mat4_inverted(const float *a, float *out) { store(load(a).inverted(), out); }
void ref_mat4_rotation(float angle, const float *axis3, float *out) { store(Matrix4::rotation(Rad(angle), Vector3{axis3[0], axis3[1], axis3[2]}), out); }
void ref_mat4_rotation_x(float angle, float *out) { store(Matrix4::rotationX(Rad(angle)), out); }
void ref_mat4_rotation_y(float angle, float *out) { store(Matrix4::rotationY(Rad(angle)), out); }
void ref_mat4_scaling_of(const float *a, float *out3) { const Vector3 s = load(a).scaling(); out3[0] = s.x(); out3[1] = s.y(); out3[2] = s.z(); }
void ref_mat4_transform_point(const float *a, const float *p3, float *out3) {
    const Vector3 r = load(a).transformPoint({p3[0], p3[1], p3[2]});
    out3[0] = r.x(); out3[1] = r.y(); out3[2] = r.z();
}
void ref_vec3_normalized(const float *v3, float *out3) {
    const Vector3 r = Vector3{v3[0], v3[1], v3[2]}.normalized();
    out3[0] = r.x(); out3[1] = r.y(); out3[2] = r.z();
}
// the reference's voxel hash (voxel_grid.hpp:39-49) and point -> voxel (voxel_grid.hpp:18-21)
unsigned long long ref_voxel_hash(int x, int y, int z) { return std::hash<Megaverse::VoxelCoords>{}(Megaverse::VoxelCoords{x, y, z}); }
void ref_to_voxel(float x, float y, float z, int *out3) {
    const Megaverse::VoxelCoords v = Megaverse::toVoxel({x, y, z});
    out3[0] = v.x(); out3[1] = v.y(); out3[2] = v.z();
}
// the reference's RNG helpers (util.hpp:25-49): n draws of randRange(lo,hi) then n of frand from mt19937(seed)
void ref_rng_stream(unsigned seed, int lo, int hi, int n, int *ints, float *floats) {
    Megaverse::Rng rng(seed);
    for (int i = 0; i < n; ++i) ints[i] = Megaverse::randRange(lo, hi, rng);
    for (int i = 0; i < n; ++i) floats[i] = Megaverse::frand(rng);
}
// iteration order of the reference's VoxelGrid hash map after inserting the given coords (component_voxel_grid.hpp:114)
int ref_voxel_grid_order(const int *xyz, int n, int *out_xyz) {
    Megaverse::VoxelGrid<int> grid(100, {0, 0, 0}, 1);
    for (int i = 0; i < n; ++i) grid.set({xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2]}, i);
    int k = 0;
    const auto copy = grid.getHashMap();
    for (auto &kv : copy) { out_xyz[k * 3] = kv.first.x(); out_xyz[k * 3 + 1] = kv.first.y(); out_xyz[k * 3 + 2] = kv.first.z(); ++k; }
    return k;
}

// Magnum SceneGraph, the way the reference's scenarios use it (Object3D = Object<MatrixTransformation3D>, util/magnum.hpp):
//   parent:  scale(ps).rotateY(angle).translate(pt)                 (component_hexagonal_maze.cpp:117)
//   child of parent:  scaleLocal(cs).translate(ct)                   (:110)
//   free object:  scale(fs).translate(ft), then setParentKeepTransformation(parent-like root = scale(rs).translate(rt))  (layout_utils.cpp:100-112)
// out: 3 x 16 floats: child's absoluteTransformationMatrix(), the same after SceneGraph::Object::setClean (the renderer's path,
// v4r_env_renderer.cpp:319-335), and the re-parented object's absoluteTransformationMatrix()
void ref_scenegraph_case(const float *ps, float angle, const float *pt, const float *cs, const float *ct, const float *fs, const float *ft,
                         const float *rs, const float *rt, float *out48) {
    using Object3D = SceneGraph::Object<SceneGraph::MatrixTransformation3D>;
    using Scene3D = SceneGraph::Scene<SceneGraph::MatrixTransformation3D>;
    Scene3D scene;
    auto &parent = scene.addChild<Object3D>();
    auto &child = parent.addChild<Object3D>();
    child.scaleLocal({cs[0], cs[1], cs[2]}).translate({ct[0], ct[1], ct[2]});
    parent.scale({ps[0], ps[1], ps[2]}).rotateY(Rad(angle)).translate({pt[0], pt[1], pt[2]});
    store(child.absoluteTransformationMatrix(), out48);
    scene.setClean();
    std::vector<std::reference_wrapper<Object3D>> objs{child};
    Object3D::setClean(objs);
    store(child.absoluteTransformationMatrix(), out48 + 16);
    auto &root = scene.addChild<Object3D>();
    root.scale({rs[0], rs[1], rs[2]}).translate({rt[0], rt[1], rt[2]});
    auto &free = scene.addChild<Object3D>();
    free.scale({fs[0], fs[1], fs[2]}).translate({ft[0], ft[1], ft[2]});
    free.setParentKeepTransformation(&root);
    store(free.absoluteTransformationMatrix(), out48 + 32);
}
// the reference's obstacle-course platforms (platforms.hpp:137-559): one platform of the given type under an anchor object at (3,1,2),
// init -> optional rotateCCW / rotateCW -> generate, then everything a scenario reads from it.  params9: obstaclesMin/MaxGap,
// Min/MaxLava, Min/MaxHeight, verticalLookLimitRad, episodeLengthSec, obstaclesNumAllowedMaxDifficulty
int ref_platform_case(int type, unsigned seed, int walls, int w, int l, int rotate, int prevWidth, const float *params9, int nObjects, int nAgents, int32_t *out, int cap) {
    using namespace Megaverse;
    using Scene3D = SceneGraph::Scene<SceneGraph::MatrixTransformation3D>;
    FloatParams fp{{"obstaclesMinGap", params9[0]}, {"obstaclesMaxGap", params9[1]}, {"obstaclesMinLava", params9[2]}, {"obstaclesMaxLava", params9[3]},
                   {"obstaclesMinHeight", params9[4]}, {"obstaclesMaxHeight", params9[5]}, {"verticalLookLimitRad", params9[6]}, {"episodeLengthSec", params9[7]},
                   {"obstaclesNumAllowedMaxDifficulty", params9[8]}};
    Rng rng(seed);
    Scene3D scene;
    auto &anchor = scene.addChild<Object3D>();
    anchor.translate({3, 1, 2});
    std::unique_ptr<Platform> p;
    switch (type) {
        case 0: p = std::make_unique<EmptyPlatform>(&anchor, rng, walls, fp, w); break;
        case 1: p = std::make_unique<WallPlatform>(&anchor, rng, walls, fp, w); break;
        case 2: p = std::make_unique<LavaPlatform>(&anchor, rng, walls, fp, w); break;
        case 3: p = std::make_unique<StepPlatform>(&anchor, rng, walls, fp, w); break;
        case 4: p = std::make_unique<GapPlatform>(&anchor, rng, walls, fp, w); break;
        case 5: p = std::make_unique<StartPlatform>(&anchor, rng, fp, w); break;
        case 6: p = std::make_unique<ExitPlatform>(&anchor, rng, fp, w); break;
        default: p = std::make_unique<TransitionPlatform>(&anchor, rng, walls, fp, l, w); break;
    }
    p->init();
    if (rotate == 1) p->rotateCCW(prevWidth); else if (rotate == 2) p->rotateCW(prevWidth);
    p->generate();
#define BX(v) (v).x()
#define BY(v) (v).y()
#define BZ(v) (v).z()

    std::vector<int32_t> o;
    auto pushBox = [&](const BoundingBox &b) { for (int v : {BX(b.min), BY(b.min), BZ(b.min), BX(b.max), BY(b.max), BZ(b.max)}) o.push_back(v); };
    for (int v : {p->length, p->height, p->width, int(p->isMaxDifficulty()), p->requiresMovableBoxesToTraverse()}) o.push_back(v);
    o.push_back(int(p->layoutBoxes.size()));
    for (auto &b : p->layoutBoxes) pushBox(b.boundingBox());
    o.push_back(int(p->wallBoxes.size()));
    for (auto &b : p->wallBoxes) pushBox(b.boundingBox());
    o.push_back(int(p->terrainBoxes.size()));
    for (auto &kv : p->terrainBoxes) { o.push_back(int(kv.first)); o.push_back(int(kv.second.size())); for (auto &b : kv.second) pushBox(b.boundingBox()); }
    pushBox(p->platformBoundingBox());

    const auto objs = p->generateObjectPositions(nObjects);
    o.push_back(int(objs.size()));
    for (auto &c : objs) for (int v : {c.x(), c.y(), c.z()}) o.push_back(v);
    const auto spawns = p->agentSpawnPoints(nAgents);
    o.push_back(int(spawns.size()));
    for (auto &c : spawns) for (float f : {c.x(), c.y(), c.z()}) { int32_t u; std::memcpy(&u, &f, 4); o.push_back(u); }
    if (p->nextPlatformAnchor) {
        const Vector3 t = p->nextPlatformAnchor->absoluteTransformation().translation();
        for (float f : {t.x(), t.y(), t.z()}) { int32_t u; std::memcpy(&u, &f, 4); o.push_back(u); }
    }
#undef BX
#undef BY
#undef BZ
    if (int(o.size()) > cap) return -int(o.size());
    std::copy(o.begin(), o.end(), out);
    return int(o.size());
}
// the reference's layout pipeline for two chained platforms: Start, then one of the obstacle types hung on its anchor (turned or
// not), both added to a VoxelGridComponent (component_voxel_grid.hpp:73-106) and merged into boxes by toBoundingBoxes (:108-187).
// out: number of (type, colour) groups, then per group type, colour, count and the boxes (inclusive min/max), in std::map order
int ref_voxel_layout_case(int type, unsigned seed, int rotate, int drawWalls, const float *params9, int32_t *out, int cap) {
    using namespace Megaverse;
    using Scene3D = SceneGraph::Scene<SceneGraph::MatrixTransformation3D>;
    FloatParams fp{{"obstaclesMinGap", params9[0]}, {"obstaclesMaxGap", params9[1]}, {"obstaclesMinLava", params9[2]}, {"obstaclesMaxLava", params9[3]},
                   {"obstaclesMinHeight", params9[4]}, {"obstaclesMaxHeight", params9[5]}, {"verticalLookLimitRad", params9[6]}, {"episodeLengthSec", params9[7]},
                   {"obstaclesNumAllowedMaxDifficulty", params9[8]}};
    Rng rng(seed);
    Scene3D scene;
    auto &levelRoot = scene.addChild<Object3D>();
    StartPlatform start(&levelRoot, rng, fp);
    start.init(); start.generate();
    std::unique_ptr<Platform> p;
    const int walls = 4 | 8, w = rotate ? -1 : start.width;
    switch (type) {
        case 1: p = std::make_unique<WallPlatform>(start.nextPlatformAnchor, rng, walls, fp, w); break;
        case 2: p = std::make_unique<LavaPlatform>(start.nextPlatformAnchor, rng, walls, fp, w); break;
        case 3: p = std::make_unique<StepPlatform>(start.nextPlatformAnchor, rng, walls, fp, w); break;
        case 4: p = std::make_unique<GapPlatform>(start.nextPlatformAnchor, rng, walls, fp, w); break;
        default: p = std::make_unique<ExitPlatform>(start.nextPlatformAnchor, rng, fp, w); break;
    }
    p->init();
    if (rotate == 1) p->rotateCCW(start.width); else if (rotate == 2) p->rotateCW(start.width);
    p->generate();
    Scenario scenario;
    VoxelGridComponent<VoxelState> vg{scenario, 100, 0, 0, 0, 1};
    vg.addPlatform(start, ColorRgb::LAYOUT_DEFAULT, ColorRgb::DARK_GREY, bool(drawWalls));
    vg.addPlatform(*p, ColorRgb::VERY_LIGHT_BLUE, ColorRgb::GREY, bool(drawWalls));
    const auto byType = vg.toBoundingBoxes();
#define BX(v) (v).x()
#define BY(v) (v).y()
#define BZ(v) (v).z()

    std::vector<int32_t> o;
    o.push_back(int(byType.size()));
    for (auto &kv : byType) {
        o.push_back(int(kv.first.type)); o.push_back(int(kv.first.color)); o.push_back(int(kv.second.size()));
        for (auto &b : kv.second) for (int v : {BX(b.min), BY(b.min), BZ(b.min), BX(b.max), BY(b.max), BZ(b.max)}) o.push_back(v);
    }
    if (int(o.size()) > cap) return -int(o.size());
    std::copy(o.begin(), o.end(), out);
    return int(o.size());

#undef BX
#undef BY
#undef BZ
}
// the reference's ObjectStackingComponent::onInteractAction (component_object_stacking.hpp:58-168) on a scripted scene:
//   solid voxels, movable 0.39 boxes sitting in voxels, agents whose object / camera matrices the script sets before every interact.
// script: per event 1 + 16 + 16 floats (agent index, agent matrix, camera-local matrix).  out per event: per object 3 translation bits
// (absolute), 3 local scaling bits, parent (-1 scene, else agent), collides; then carrying[agent] per agent; then the voxels holding
// an object as sorted (x, y, z, object) rows, preceded by their count.
int ref_stacking_case(const int *solid, int nSolid, const int *objVoxels, int nObj, int nAgents, const float *script, int nEvents, int32_t *out, int cap) {
    using namespace Megaverse;
    struct TestAgent : AbstractAgent {
        explicit TestAgent(Object3D *parent) : AbstractAgent{parent} {
            cameraObject = &addChild<Object3D>();
            pickupSpot = &cameraObject->addChild<Object3D>();
            pickupSpot->translate({0.0f, -0.44f, -1.0f});  // agent.cpp:40
        }
        Object3D *interactLocation() override { return pickupSpot; }
        Object3D *cameraObject, *pickupSpot;
    };
    struct NoCallbacks : ObjectStackingCallbacks {} callbacks;
    Env env;
    env.numAgents = nAgents;
    Env::EnvState st;
    st.scene = std::make_unique<Scene3D>();
    st.physics = std::make_unique<Env::Physics>();
    VoxelGrid<VoxelWithPhysicsObjects> grid(100, {0, 0, 0}, 1);
    for (int i = 0; i < nSolid; ++i) grid.set({solid[i * 3], solid[i * 3 + 1], solid[i * 3 + 2]}, makeVoxel<VoxelWithPhysicsObjects>(VOXEL_SOLID | VOXEL_OPAQUE));
    std::vector<RigidBody *> objects;
    btBoxShape shape(btVector3{1, 1, 1});
    for (int i = 0; i < nObj; ++i) {  // what addDrawablesAndCollisions does per object (:170-198), minus Bullet and the drawable
        const VoxelCoords pos{objVoxels[i * 3], objVoxels[i * 3 + 1], objVoxels[i * 3 + 2]};
        auto &object = st.scene->addChild<RigidBody>(st.scene.get(), 0.0f, &shape, st.physics->bWorld);
        object.scale({0.39f, 0.39f, 0.39f}).translate({float(pos.x()) + 0.5f, float(pos.y()) + 0.5f, float(pos.z()) + 0.5f});
        if (!grid.hasVoxel(pos)) grid.set(pos, VoxelWithPhysicsObjects{});
        grid.get(pos)->physicsObject = &object;
        objects.push_back(&object);
    }
    std::vector<TestAgent *> agents;
    for (int i = 0; i < nAgents; ++i) { agents.push_back(&st.scene->addChild<TestAgent>(st.scene.get())); st.agents.push_back(agents.back()); }
    Scenario scenario;
    ObjectStackingComponent<VoxelWithPhysicsObjects> stacking{scenario, nAgents, grid, callbacks};
    stacking.reset(env, st);
    std::vector<int32_t> o;
    auto bits = [](float f) { int32_t u; std::memcpy(&u, &f, 4); return u; };
    for (int ev = 0; ev < nEvents; ++ev) {
        const float *e = script + size_t(ev) * 33;
        const int ai = int(e[0]);
        agents[size_t(ai)]->setTransformation(Matrix4::from(e + 1));
        agents[size_t(ai)]->cameraObject->setTransformation(Matrix4::from(e + 17));
        stacking.onInteractAction(ai, st);
        for (auto *obj : objects) {
            const Vector3 t = obj->absoluteTransformation().translation(), sc = obj->transformation().scaling();
            for (float f : {t.x(), t.y(), t.z(), sc.x(), sc.y(), sc.z()}) o.push_back(bits(f));
            int parent = -1;
            for (int a = 0; a < nAgents; ++a) if (obj->parent() == agents[size_t(a)]->pickupSpot) parent = a;
            o.push_back(parent); o.push_back(obj->colliding() ? 1 : 0);
        }
        for (int a = 0; a < nAgents; ++a) {
            int carried = -1;
            for (int k = 0; k < nObj; ++k) if (stacking.agentCarryingObject(a) == objects[size_t(k)]) carried = k;
            o.push_back(carried);
        }
        std::vector<std::array<int32_t, 4>> occ;
        for (auto &kv : grid.getHashMap())
            if (kv.second.physicsObject)
                for (int k = 0; k < nObj; ++k) if (kv.second.physicsObject == objects[size_t(k)]) occ.push_back({kv.first.x(), kv.first.y(), kv.first.z(), k});
        std::sort(occ.begin(), occ.end());
        o.push_back(int(occ.size()));
        for (auto &r : occ) for (int v : r) o.push_back(v);
    }
    if (int(o.size()) > cap) return -int(o.size());
    std::copy(o.begin(), o.end(), out);
    return int(o.size());
}
// the reference's layout_utils.cpp (compiled in place): what each helper adds to the drawables map and to the collision world.
// out: per drawable type 0..4 its count, then per drawable 16 matrix bits (absoluteTransformationMatrix) + colour 0xRRGGBB-as-floats
// (3 bits); then the number of rigid bodies and per body 6 bits (collider origin, collider scaling)
int ref_layout_utils_case(unsigned seed, int32_t *out, int cap) {
    using namespace Megaverse;
    Rng rng(seed);
    Env::EnvState st;
    st.scene = std::make_unique<Scene3D>();
    st.physics = std::make_unique<Env::Physics>();
    DrawablesMap drawables;
    auto fr = [&](float lo, float hi) { return lo + (hi - lo) * frand(rng); };
    Boxes boxes;
    for (int i = 0; i < 5; ++i) {
        const int x = randRange(-5, 20, rng), y = randRange(0, 4, rng), z = randRange(-5, 20, rng);
        const int x1 = x + randRange(0, 9, rng), y1 = y + randRange(0, 3, rng), z1 = z + randRange(0, 9, rng);  // (named: argument order is unspecified)
        boxes.emplace_back(x, y, z, x1, y1, z1);
    }
    addBoundingBoxes(drawables, st, boxes, VOXEL_SOLID | VOXEL_OPAQUE, ColorRgb::GREY, 1.0f);
    addBoundingBoxes(drawables, st, boxes, VOXEL_SOLID, ColorRgb::DARK_GREY, 2.0f);   // invisible colliders, voxel size 2 (Sokoban)
    addBoundingBoxes(drawables, st, boxes, VOXEL_OPAQUE, ColorRgb::ORANGE, 1.0f);     // drawn, no collider
    for (int i = 0; i < 4; ++i) {
        const int x = randRange(0, 20, rng), z = randRange(0, 20, rng);
        const int y = randRange(1, 3, rng), x1 = x + randRange(0, 6, rng), z1 = z + randRange(1, 6, rng);
        addTerrain(drawables, st, i % 2 ? TERRAIN_LAVA : TERRAIN_EXIT, BoundingBox{x, y, z, x1, 2, z1}, 1.0f);
    }
    for (int i = 0; i < 4; ++i) {
        const float sx = fr(0.1f, 9), sy = fr(0.0001f, 2), sz = fr(0.1f, 9);
        const float tx = fr(-20, 20), ty = fr(-1, 3), tz = fr(-20, 20);
        addStaticCollidingBox(drawables, st, {sx, sy, sz}, {tx, ty, tz}, ColorRgb::BLUE);
    }
    for (int i = 0; i < 3; ++i) {
        { const float x = fr(-20, 20), y = fr(0, 3), z = fr(-20, 20); const float k = fr(0.5f, 2.5f); addDiamond(drawables, *st.scene, {x, y, z}, Magnum::Vector3{0.17f, 0.45f, 0.17f} * k, ColorRgb::GREEN); }
        { const float x = fr(-20, 20), y = fr(0, 3), z = fr(-20, 20); const float k = fr(0.5f, 1.5f); addPillar(drawables, *st.scene, {x, y, z}, Magnum::Vector3{0.5f, 2, 0.5f} * k, ColorRgb::VIOLET); }
        { const float x = fr(-20, 20), y = fr(0, 3), z = fr(-20, 20); const float k = fr(0.5f, 1.5f); addSphere(drawables, *st.scene, {x, y, z}, Magnum::Vector3{0.75f, 0.75f, 0.75f} * k, ColorRgb::RED); }
    }
    std::vector<int32_t> o;
    auto bits = [](float f) { int32_t u; std::memcpy(&u, &f, 4); return u; };
    for (int t = 0; t < 5; ++t) o.push_back(int(drawables[DrawableType(t)].size()));
    for (int t = 0; t < 5; ++t)
        for (auto &d : drawables[DrawableType(t)]) {
            const Matrix4 m = d.objectPtr->absoluteTransformationMatrix();
            for (int i = 0; i < 16; ++i) o.push_back(bits(m.data()[i]));
            for (float f : {d.color.r(), d.color.g(), d.color.b()}) o.push_back(bits(f));
        }
    std::vector<RigidBody *> bodies;
    std::function<void(Object3D &)> walk = [&](Object3D &n) {
        if (auto *rb = dynamic_cast<RigidBody *>(&n)) bodies.push_back(rb);
        for (Object3D &c : n.children()) walk(c);
    };
    // creation order: children lists are in insertion order
    for (Object3D &c : st.scene->children()) walk(c);
    o.push_back(int(bodies.size()));
    for (auto *rb : bodies)
        for (float f : {rb->colliderOrigin.x(), rb->colliderOrigin.y(), rb->colliderOrigin.z(), rb->colliderScaling.x(), rb->colliderScaling.y(), rb->colliderScaling.z()}) o.push_back(bits(f));
    if (int(o.size()) > cap) return -int(o.size());
    std::copy(o.begin(), o.end(), out);
    return int(o.size());
}
// the reference's colour tables (env/const.hpp:25-143): [n all, n agent, n object, n layout] then the 0xRRGGBB values in that order,
// then rgb(allColors[i]) as 3 floats each (bit patterns), i.e. what the renderer multiplies with
int ref_color_tables(unsigned *out, int cap) {
    using namespace Megaverse;
    std::vector<unsigned> o{unsigned(numColors), unsigned(numAgentColors), unsigned(numObjectColors), unsigned(numLayoutColors)};
    for (int i = 0; i < numColors; ++i) o.push_back(unsigned(allColors[i]));
    for (int i = 0; i < numAgentColors; ++i) o.push_back(unsigned(agentColors[i]));
    for (int i = 0; i < numObjectColors; ++i) o.push_back(unsigned(objectColors[i]));
    for (int i = 0; i < numLayoutColors; ++i) o.push_back(unsigned(layoutColors[i]));
    for (int i = 0; i < numColors; ++i) {
        const Magnum::Color3 c = rgb(allColors[i]);
        for (float f : {c.r(), c.g(), c.b()}) { unsigned u; std::memcpy(&u, &f, 4); o.push_back(u); }
    }
    if (int(o.size()) > cap) return -int(o.size());
    std::copy(o.begin(), o.end(), out);
    return int(o.size());
}
// the reference's Perlin noise (scenario_collect.cpp:57-75 uses accumulatedOctaveNoise2D_0_1 on a PerlinNoise(seed))
void ref_perlin(unsigned seed, int n, const double *xy, int octaves, double *out) {
    const siv::PerlinNoise perlin(seed);
    for (int i = 0; i < n; ++i) out[i] = perlin.accumulatedOctaveNoise2D_0_1(xy[2 * i], xy[2 * i + 1], octaves);
}
// the reference's honeycomb maze (src/libs/mazes) with the Kruskal generator seeded explicitly (upstream: random_device).
// out: [vertices, then per cell: centre x, centre y, number of adjacency entries, then per entry: neighbour, x1, y1, x2, y2],
// followed by the four coordinate bounds
int ref_honeycomb_maze(int size, unsigned seed, double *out, int cap) {
    struct SeededKruskal : Kruskal { explicit SeededKruskal(unsigned s) { generator = std::mt19937(s); } };
    HoneyCombMaze maze(size);
    SeededKruskal algorithm(seed);
    maze.InitialiseGraph();
    maze.GenerateMaze(&algorithm);
    std::vector<double> o;
    auto &adj = maze.getAdjacencyList();
    o.push_back(double(adj.size()));
    for (size_t c = 0; c < adj.size(); ++c) {
        o.push_back(maze.getCellCenters()[c].first); o.push_back(maze.getCellCenters()[c].second); o.push_back(double(adj[c].size()));
        for (auto &e : adj[c]) {
            const auto [x1, y1, x2, y2] = dynamic_cast<LineBorder *>(e.second.get())->getBorderCoords();
            for (double v : {double(e.first), x1, y1, x2, y2}) o.push_back(v);
        }
    }
    const auto [xmin, ymin, xmax, ymax] = maze.GetCoordinateBounds();
    for (double v : {xmin, ymin, xmax, ymax}) o.push_back(v);
    if (int(o.size()) > cap) return -int(o.size());
    std::copy(o.begin(), o.end(), out);
    return int(o.size());
}

}  // extern "C"
