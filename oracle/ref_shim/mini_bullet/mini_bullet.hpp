// ORACLE tooling -- NOT product code, and NOT Bullet.  A stand-in for the slice of the Bullet 2.89 API that the reference's env
// library touches (env/physics.hpp, env/agent.hpp/.cpp, env/env.hpp/.cpp, env/kinematic_character_controller.hpp/.cpp, the scenario
// sources and magnum-integration's BulletIntegration glue), so that ALL of that code compiles unmodified, in place, without Bullet
// (which is neither vendored under /root/reference nor installed here).  What is real and what is stand-in:
//   * real, compiled from /root/reference: the character controller's state machine, the agent, Env::reset / Env::step, RigidBody
//     and its syncPose, MotionState, every scenario and component;
//   * stand-in, this file: (1) LinearMath -- btVector3 / btQuaternion / btMatrix3x3 / btTransform with the arithmetic the oracle
//     restates in oracle/orc_math.hpp (operation order of Bullet's scalar code paths; transcendental functions as the oracle defines
//     them), (2) the collision world's containers and the order it visits objects in, (3) the NARROW PHASE: convexSweepTest and
//     the contact manifolds are answered by the oracle's analytic definitions (oracle/orc_physics.hpp: sweepBroadphaseMiss,
//     sweepNarrow, capsuleDistance) instead of Bullet's GJK / conservative advancement.
// So a run of the reference on this stand-in pins everything the oracle restates ABOVE the narrow phase -- the part it cannot pin
// is exactly the part orc_physics.hpp's header declares as replaced.
//
// Visiting order (a modelling choice the oracle shares, Bullet's own order depends on its broadphase internals): static bodies in
// the order they joined the world, then character (ghost) objects in the order they joined.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "../../orc_physics.hpp"

typedef float btScalar;
#define SIMD_EPSILON FLT_EPSILON
#define SIMD_PI btScalar(3.1415926535897932384626433832795029)
#define SIMD_2_PI (btScalar(2.0) * SIMD_PI)
#define SIMD_HALF_PI (SIMD_PI * btScalar(0.5))
#define SIMD_RADS_PER_DEG (SIMD_2_PI / btScalar(360.0))
#define ATTRIBUTE_ALIGNED16(a) a
#define BT_DECLARE_ALIGNED_ALLOCATOR()
#define btAssert(x)
#define DISABLE_DEACTIVATION 4
#define DISABLE_SIMULATION 5

inline btScalar btSqrt(btScalar x) { return sqrtf(x); }
inline btScalar btFabs(btScalar x) { return fabsf(x); }
inline btScalar btCos(btScalar x) { return orc::crcos(x); }
inline btScalar btSin(btScalar x) { return orc::crsin(x); }
inline btScalar btAcos(btScalar x) {  // btScalar.h: clamps to [-1, 1]
    if (x < btScalar(-1)) x = btScalar(-1);
    if (x > btScalar(1)) x = btScalar(1);
    return orc::cracos(x);
}
inline btScalar btPow(btScalar x, btScalar y) { return powf(x, y); }
inline btScalar btRadians(btScalar x) { return x * SIMD_RADS_PER_DEG; }
template <typename T> inline const T &btClamped(const T &a, const T &lb, const T &ub) { return a < lb ? lb : (ub < a ? ub : a); }

// ---------------------------------------------------------------- LinearMath
class btVector3 {
public:
    btScalar m_floats[4];
    btVector3() {}
    btVector3(const btScalar &x, const btScalar &y, const btScalar &z) { m_floats[0] = x, m_floats[1] = y, m_floats[2] = z, m_floats[3] = 0; }
    const btScalar &x() const { return m_floats[0]; }
    const btScalar &y() const { return m_floats[1]; }
    const btScalar &z() const { return m_floats[2]; }
    const btScalar &getX() const { return m_floats[0]; }
    const btScalar &getY() const { return m_floats[1]; }
    const btScalar &getZ() const { return m_floats[2]; }
    void setX(btScalar v) { m_floats[0] = v; }
    void setY(btScalar v) { m_floats[1] = v; }
    void setZ(btScalar v) { m_floats[2] = v; }
    void setValue(const btScalar &x, const btScalar &y, const btScalar &z) { m_floats[0] = x, m_floats[1] = y, m_floats[2] = z, m_floats[3] = 0; }
    void setZero() { setValue(0, 0, 0); }
    operator btScalar *() { return &m_floats[0]; }
    operator const btScalar *() const { return &m_floats[0]; }
    btVector3 &operator+=(const btVector3 &v) { m_floats[0] += v.m_floats[0], m_floats[1] += v.m_floats[1], m_floats[2] += v.m_floats[2]; return *this; }
    btVector3 &operator-=(const btVector3 &v) { m_floats[0] -= v.m_floats[0], m_floats[1] -= v.m_floats[1], m_floats[2] -= v.m_floats[2]; return *this; }
    btVector3 &operator*=(const btScalar &s) { m_floats[0] *= s, m_floats[1] *= s, m_floats[2] *= s; return *this; }
    btVector3 &operator/=(const btScalar &s) { return *this *= btScalar(1.0) / s; }
    btScalar dot(const btVector3 &v) const { return m_floats[0] * v.m_floats[0] + m_floats[1] * v.m_floats[1] + m_floats[2] * v.m_floats[2]; }
    btScalar length2() const { return dot(*this); }
    btScalar length() const { return btSqrt(length2()); }
    btVector3 &normalize() { return *this /= length(); }
    btVector3 normalized() const { btVector3 n = *this; return n.normalize(); }
    bool fuzzyZero() const { return length2() < SIMD_EPSILON * SIMD_EPSILON; }
    btVector3 cross(const btVector3 &v) const {
        return btVector3(m_floats[1] * v.m_floats[2] - m_floats[2] * v.m_floats[1], m_floats[2] * v.m_floats[0] - m_floats[0] * v.m_floats[2],
                         m_floats[0] * v.m_floats[1] - m_floats[1] * v.m_floats[0]);
    }
    void setInterpolate3(const btVector3 &v0, const btVector3 &v1, btScalar rt) {
        const btScalar s = btScalar(1.0) - rt;
        m_floats[0] = s * v0.m_floats[0] + rt * v1.m_floats[0];
        m_floats[1] = s * v0.m_floats[1] + rt * v1.m_floats[1];
        m_floats[2] = s * v0.m_floats[2] + rt * v1.m_floats[2];
    }
    bool operator==(const btVector3 &o) const { return m_floats[0] == o.m_floats[0] && m_floats[1] == o.m_floats[1] && m_floats[2] == o.m_floats[2] && m_floats[3] == o.m_floats[3]; }
    bool operator!=(const btVector3 &o) const { return !(*this == o); }
};
inline btVector3 operator+(const btVector3 &a, const btVector3 &b) { return btVector3(a.m_floats[0] + b.m_floats[0], a.m_floats[1] + b.m_floats[1], a.m_floats[2] + b.m_floats[2]); }
inline btVector3 operator-(const btVector3 &a, const btVector3 &b) { return btVector3(a.m_floats[0] - b.m_floats[0], a.m_floats[1] - b.m_floats[1], a.m_floats[2] - b.m_floats[2]); }
inline btVector3 operator-(const btVector3 &a) { return btVector3(-a.m_floats[0], -a.m_floats[1], -a.m_floats[2]); }
inline btVector3 operator*(const btVector3 &a, const btVector3 &b) { return btVector3(a.m_floats[0] * b.m_floats[0], a.m_floats[1] * b.m_floats[1], a.m_floats[2] * b.m_floats[2]); }
inline btVector3 operator*(const btVector3 &a, const btScalar &s) { return btVector3(a.m_floats[0] * s, a.m_floats[1] * s, a.m_floats[2] * s); }
inline btVector3 operator*(const btScalar &s, const btVector3 &a) { return a * s; }
inline btVector3 operator/(const btVector3 &a, const btScalar &s) { return a * (btScalar(1.0) / s); }

inline orc::Vec3 toOrc(const btVector3 &v) { return {v.x(), v.y(), v.z()}; }
inline btVector3 fromOrc(const orc::Vec3 &v) { return btVector3(v.x, v.y, v.z); }

class btQuaternion {
public:
    btScalar m_floats[4];
    btQuaternion() {}
    btQuaternion(const btScalar &x, const btScalar &y, const btScalar &z, const btScalar &w) { m_floats[0] = x, m_floats[1] = y, m_floats[2] = z, m_floats[3] = w; }
    btQuaternion(const btVector3 &axis, const btScalar &angle) { setRotation(axis, angle); }
    void setRotation(const btVector3 &axis, const btScalar &angle) {
        const orc::Quat q = orc::quatAxisAngle(toOrc(axis), angle);
        m_floats[0] = q.x, m_floats[1] = q.y, m_floats[2] = q.z, m_floats[3] = q.w;
    }
    const btScalar &x() const { return m_floats[0]; }
    const btScalar &y() const { return m_floats[1]; }
    const btScalar &z() const { return m_floats[2]; }
    const btScalar &w() const { return m_floats[3]; }
    orc::Quat orcQ() const { return {m_floats[0], m_floats[1], m_floats[2], m_floats[3]}; }
    btScalar getAngle() const { return orc::quatAngle(orcQ()); }
    btVector3 getAxis() const { return fromOrc(orc::quatAxis(orcQ())); }
    btQuaternion inverse() const { return btQuaternion(-m_floats[0], -m_floats[1], -m_floats[2], m_floats[3]); }
    btScalar length2() const { return m_floats[0] * m_floats[0] + m_floats[1] * m_floats[1] + m_floats[2] * m_floats[2] + m_floats[3] * m_floats[3]; }
    bool operator==(const btQuaternion &o) const { return !std::memcmp(m_floats, o.m_floats, sizeof(m_floats)) || (m_floats[0] == o.m_floats[0] && m_floats[1] == o.m_floats[1] && m_floats[2] == o.m_floats[2] && m_floats[3] == o.m_floats[3]); }
};
inline btQuaternion operator*(const btQuaternion &q1, const btQuaternion &q2) {
    return btQuaternion(q1.w() * q2.x() + q1.x() * q2.w() + q1.y() * q2.z() - q1.z() * q2.y(), q1.w() * q2.y() + q1.y() * q2.w() + q1.z() * q2.x() - q1.x() * q2.z(),
                        q1.w() * q2.z() + q1.z() * q2.w() + q1.x() * q2.y() - q1.y() * q2.x(), q1.w() * q2.w() - q1.x() * q2.x() - q1.y() * q2.y() - q1.z() * q2.z());
}
inline btQuaternion shortestArcQuat(const btVector3 &v0, const btVector3 &v1) {  // btQuaternion.h
    const btVector3 c = v0.cross(v1);
    const btScalar d = v0.dot(v1);
    if (d < -1.0 + SIMD_EPSILON) return btQuaternion(0.0f, 1.0f, 0.0f, 0.0f);  // (any perpendicular axis; unreachable here: up stays +Y)
    const btScalar s = btSqrt((1.0f + d) * 2.0f);
    const btScalar rs = 1.0f / s;
    return btQuaternion(c.getX() * rs, c.getY() * rs, c.getZ() * rs, s * 0.5f);
}
inline btQuaternion shortestArcQuatNormalize2(btVector3 &v0, btVector3 &v1) {
    v0.normalize();
    v1.normalize();
    return shortestArcQuat(v0, v1);
}

class btMatrix3x3 {
public:
    btVector3 m_el[3];
    btMatrix3x3() {}
    explicit btMatrix3x3(const btQuaternion &q) { setRotation(q); }
    btMatrix3x3(const btScalar &xx, const btScalar &xy, const btScalar &xz, const btScalar &yx, const btScalar &yy, const btScalar &yz, const btScalar &zx, const btScalar &zy,
                const btScalar &zz) {
        m_el[0].setValue(xx, xy, xz), m_el[1].setValue(yx, yy, yz), m_el[2].setValue(zx, zy, zz);
    }
    static btMatrix3x3 fromOrc(const orc::Mat3 &m) { return btMatrix3x3(m.r[0].x, m.r[0].y, m.r[0].z, m.r[1].x, m.r[1].y, m.r[1].z, m.r[2].x, m.r[2].y, m.r[2].z); }
    orc::Mat3 orcM() const { orc::Mat3 m; for (int i = 0; i < 3; ++i) m.r[i] = toOrc(m_el[i]); return m; }
    void setRotation(const btQuaternion &q) { *this = fromOrc(orc::mat3FromQuat(q.orcQ())); }
    void getRotation(btQuaternion &q) const { const orc::Quat o = orc::quatFromMat3(orcM()); q = btQuaternion(o.x, o.y, o.z, o.w); }
    void setIdentity() { *this = btMatrix3x3(1, 0, 0, 0, 1, 0, 0, 0, 1); }
    btVector3 getColumn(int i) const { return btVector3(m_el[0][i], m_el[1][i], m_el[2][i]); }
    const btVector3 &getRow(int i) const { return m_el[i]; }
    btVector3 &operator[](int i) { return m_el[i]; }
    const btVector3 &operator[](int i) const { return m_el[i]; }
    btMatrix3x3 &operator*=(const btMatrix3x3 &m) { *this = fromOrc(orc::mul(orcM(), m.orcM())); return *this; }
    void setFromOpenGLSubMatrix(const btScalar *m) { m_el[0].setValue(m[0], m[4], m[8]), m_el[1].setValue(m[1], m[5], m[9]), m_el[2].setValue(m[2], m[6], m[10]); }
    void getOpenGLSubMatrix(btScalar *m) const {
        m[0] = m_el[0].x(), m[1] = m_el[1].x(), m[2] = m_el[2].x(), m[3] = 0, m[4] = m_el[0].y(), m[5] = m_el[1].y(), m[6] = m_el[2].y(), m[7] = 0;
        m[8] = m_el[0].z(), m[9] = m_el[1].z(), m[10] = m_el[2].z(), m[11] = 0;
    }
    bool operator==(const btMatrix3x3 &o) const { return m_el[0] == o.m_el[0] && m_el[1] == o.m_el[1] && m_el[2] == o.m_el[2]; }
};
inline btVector3 operator*(const btMatrix3x3 &m, const btVector3 &v) { return btVector3(m[0].dot(v), m[1].dot(v), m[2].dot(v)); }
inline btMatrix3x3 operator*(const btMatrix3x3 &a, const btMatrix3x3 &b) { return btMatrix3x3::fromOrc(orc::mul(a.orcM(), b.orcM())); }

class btTransform {
public:
    btMatrix3x3 m_basis;
    btVector3 m_origin;
    btTransform() {}
    explicit btTransform(const btMatrix3x3 &b, const btVector3 &c = btVector3(0, 0, 0)) : m_basis(b), m_origin(c) {}
    explicit btTransform(const btQuaternion &q, const btVector3 &c = btVector3(0, 0, 0)) : m_basis(q), m_origin(c) {}
    void setIdentity() { m_basis.setIdentity(); m_origin.setValue(0, 0, 0); }
    btMatrix3x3 &getBasis() { return m_basis; }
    const btMatrix3x3 &getBasis() const { return m_basis; }
    btVector3 &getOrigin() { return m_origin; }
    const btVector3 &getOrigin() const { return m_origin; }
    btQuaternion getRotation() const { btQuaternion q; m_basis.getRotation(q); return q; }
    void setOrigin(const btVector3 &o) { m_origin = o; }
    void setBasis(const btMatrix3x3 &b) { m_basis = b; }
    void setRotation(const btQuaternion &q) { m_basis.setRotation(q); }
    void setFromOpenGLMatrix(const btScalar *m) { m_basis.setFromOpenGLSubMatrix(m); m_origin.setValue(m[12], m[13], m[14]); }
    void getOpenGLMatrix(btScalar *m) const { m_basis.getOpenGLSubMatrix(m); m[12] = m_origin.x(), m[13] = m_origin.y(), m[14] = m_origin.z(), m[15] = btScalar(1.0); }
    bool operator==(const btTransform &o) const { return m_basis == o.m_basis && m_origin == o.m_origin; }
};

class btIDebugDraw {};
class btMotionState {
public:
    virtual ~btMotionState() = default;
    virtual void getWorldTransform(btTransform &worldTrans) const = 0;
    virtual void setWorldTransform(const btTransform &worldTrans) = 0;
};

template <typename T> class btAlignedObjectArray {
public:
    std::vector<T> v;
    int size() const { return int(v.size()); }
    void resize(int n) { v.resize(size_t(n)); }
    void push_back(const T &t) { v.push_back(t); }
    void clear() { v.clear(); }
    T &operator[](int i) { return v[size_t(i)]; }
    const T &operator[](int i) const { return v[size_t(i)]; }
};

// ---------------------------------------------------------------- collision shapes and objects
struct btBroadphaseProxy {
    enum CollisionFilterGroups { DefaultFilter = 1, StaticFilter = 2, KinematicFilter = 4, DebrisFilter = 8, SensorTrigger = 16, CharacterFilter = 32, AllFilter = -1 };
    void *m_clientObject = nullptr;
    int m_collisionFilterGroup = 0, m_collisionFilterMask = 0;
};

class btCollisionShape {
public:
    enum Kind { BOX, CAPSULE, OTHER } kind = OTHER;
    btVector3 m_localScaling{1, 1, 1}, m_dims{0, 0, 0};  // box: unscaled half extents; capsule: (radius, cylinder height, 0)
    virtual ~btCollisionShape() = default;
    virtual void setLocalScaling(const btVector3 &s) { m_localScaling = s; }
    const btVector3 &getLocalScaling() const { return m_localScaling; }
    virtual void calculateLocalInertia(btScalar, btVector3 &inertia) const { inertia.setValue(0, 0, 0); }
};
class btConvexShape : public btCollisionShape {
public:
    void getAabb(const btTransform &, btVector3 &mn, btVector3 &mx) const { mn.setValue(0, 0, 0), mx.setValue(0, 0, 0); }  // the stand-in broadphase ignores it
};
class btBoxShape : public btConvexShape {
public:
    explicit btBoxShape(const btVector3 &halfExtents) { kind = BOX; m_dims = halfExtents; }
};
class btCapsuleShape : public btConvexShape {
public:
    btCapsuleShape(btScalar radius, btScalar height) { kind = CAPSULE; m_dims.setValue(radius, height, 0); }
};

class btCollisionWorld;
class btCollisionObject {
public:
    enum CollisionFlags { CF_STATIC_OBJECT = 1, CF_KINEMATIC_OBJECT = 2, CF_NO_CONTACT_RESPONSE = 4, CF_CUSTOM_MATERIAL_CALLBACK = 8, CF_CHARACTER_OBJECT = 16 };
    btCollisionObject() { m_worldTransform.setIdentity(); m_proxy.m_clientObject = this; }
    virtual ~btCollisionObject() = default;
    btTransform &getWorldTransform() { return m_worldTransform; }
    const btTransform &getWorldTransform() const { return m_worldTransform; }
    void setWorldTransform(const btTransform &t) { m_worldTransform = t; }
    bool hasContactResponse() const { return (m_collisionFlags & CF_NO_CONTACT_RESPONSE) == 0; }
    int getCollisionFlags() const { return m_collisionFlags; }
    void setCollisionFlags(int f) { m_collisionFlags = f; }
    btCollisionShape *getCollisionShape() { return m_shape; }
    const btCollisionShape *getCollisionShape() const { return m_shape; }
    void setCollisionShape(btCollisionShape *s) { m_shape = s; }
    btBroadphaseProxy *getBroadphaseHandle() { return &m_proxy; }
    const btBroadphaseProxy *getBroadphaseHandle() const { return &m_proxy; }
    void forceActivationState(int) const {}
    void setActivationState(int) const {}

    // what the oracle's analytic narrow phase needs to know about this object
    orc::Collider collider() const {
        orc::Collider c;
        const btTransform &t = m_worldTransform;
        c.c = toOrc(t.getOrigin());
        c.enabled = hasContactResponse();
        if (m_shape && m_shape->kind == btCollisionShape::CAPSULE) { c.kind = 1; return c; }
        c.kind = 0;
        const btVector3 h = m_shape->m_dims * m_shape->getLocalScaling();
        c.h = toOrc(h);
        // a box whose basis has no off-diagonal x/z terms is axis aligned (its diagonal may be 1 - 1ulp: s * (1/s)); the maze walls
        // are turned about Y: local x axis = first column = (ax, 0, az)
        const btMatrix3x3 &b = t.getBasis();
        if (b[2].x() != 0.0f || b[0].z() != 0.0f) { c.rotated = true; c.ax = b[0].x(); c.az = b[2].x(); }
        return c;
    }

    btTransform m_worldTransform;
    btCollisionShape *m_shape = nullptr;
    int m_collisionFlags = CF_STATIC_OBJECT;
    btBroadphaseProxy m_proxy;
    btCollisionWorld *m_world = nullptr;
};

class btRigidBody : public btCollisionObject {
public:
    struct btRigidBodyConstructionInfo {
        btRigidBodyConstructionInfo(btScalar mass, btMotionState *motionState, btCollisionShape *shape, const btVector3 &inertia = btVector3(0, 0, 0))
        : m_mass(mass), m_motionState(motionState), m_collisionShape(shape), m_localInertia(inertia) {}
        btScalar m_mass;
        btMotionState *m_motionState;
        btCollisionShape *m_collisionShape;
        btVector3 m_localInertia;
    };
    explicit btRigidBody(const btRigidBodyConstructionInfo &info) {  // btRigidBody::setupRigidBody
        m_shape = info.m_collisionShape;
        if (info.m_motionState) info.m_motionState->getWorldTransform(m_worldTransform);
    }
};

// ---------------------------------------------------------------- manifolds, pair cache
class btManifoldPoint {
public:
    btVector3 m_normalWorldOnB;
    btScalar m_distance1 = 0;
    btScalar getDistance() const { return m_distance1; }
};
class btPersistentManifold {
public:
    const btCollisionObject *m_body0 = nullptr, *m_body1 = nullptr;
    std::vector<btManifoldPoint> m_points;
    const btCollisionObject *getBody0() const { return m_body0; }
    const btCollisionObject *getBody1() const { return m_body1; }
    int getNumContacts() const { return int(m_points.size()); }
    const btManifoldPoint &getContactPoint(int i) const { return m_points[size_t(i)]; }
};
typedef btAlignedObjectArray<btPersistentManifold *> btManifoldArray;
class btCollisionAlgorithm {
public:
    btPersistentManifold m_manifold;
    void getAllContactManifolds(btManifoldArray &out) { out.push_back(&m_manifold); }
};
struct btBroadphasePair {
    btBroadphaseProxy *m_pProxy0 = nullptr, *m_pProxy1 = nullptr;
    btCollisionAlgorithm *m_algorithm = nullptr;
};
typedef btAlignedObjectArray<btBroadphasePair> btBroadphasePairArray;

class btDispatcher;
class btPairCachingGhostObject;
class btGhostPairCallback {};
class btOverlappingPairCache {
public:
    virtual ~btOverlappingPairCache() = default;
    void setInternalGhostPairCallback(btGhostPairCallback *) {}
};
class btHashedOverlappingPairCache : public btOverlappingPairCache {
public:
    btPairCachingGhostObject *m_owner = nullptr;
    btBroadphasePairArray m_pairs;
    std::vector<std::unique_ptr<btCollisionAlgorithm>> m_algorithms;
    int getNumOverlappingPairs() const { return m_pairs.size(); }
    btBroadphasePairArray &getOverlappingPairArray() { return m_pairs; }
    void removeOverlappingPair(btBroadphaseProxy *, btBroadphaseProxy *, btDispatcher *) { if (m_pairs.size()) m_pairs.v.erase(m_pairs.v.begin()); }
};

// ---------------------------------------------------------------- world
struct btDispatcherInfo {
    btScalar m_allowedCcdPenetration = btScalar(0.04);  // btDispatcher.h default
};
class btCollisionConfiguration {};
class btDefaultCollisionConfiguration : public btCollisionConfiguration {};
class btConstraintSolver {};
class btSequentialImpulseConstraintSolver : public btConstraintSolver {};
class btDispatcher {
public:
    // btCollisionDispatcher::dispatchAllCollisionPairs for one ghost's private pair cache: (re)compute its contacts, defined below
    void dispatchAllCollisionPairs(btOverlappingPairCache *cache, const btDispatcherInfo &, btDispatcher *);
};
class btCollisionDispatcher : public btDispatcher {
public:
    explicit btCollisionDispatcher(btCollisionConfiguration *) {}
};
class btBroadphaseInterface {
public:
    btOverlappingPairCache m_cache;
    btOverlappingPairCache *getOverlappingPairCache() { return &m_cache; }
    void setAabb(btBroadphaseProxy *, const btVector3 &, const btVector3 &, btDispatcher *) {}  // every object is a candidate (orc_physics.hpp header)
};
class btDbvtBroadphase : public btBroadphaseInterface {};

class btActionInterface {
public:
    virtual ~btActionInterface() = default;
    virtual void updateAction(btCollisionWorld *collisionWorld, btScalar deltaTimeStep) = 0;
    virtual void debugDraw(btIDebugDraw *debugDrawer) = 0;
};

class btCollisionWorld {
public:
    struct LocalShapeInfo {};
    struct LocalConvexResult {
        LocalConvexResult(const btCollisionObject *o, LocalShapeInfo *s, const btVector3 &n, const btVector3 &p, btScalar f)
        : m_hitCollisionObject(o), m_localShapeInfo(s), m_hitNormalLocal(n), m_hitPointLocal(p), m_hitFraction(f) {}
        const btCollisionObject *m_hitCollisionObject;
        LocalShapeInfo *m_localShapeInfo;
        btVector3 m_hitNormalLocal, m_hitPointLocal;
        btScalar m_hitFraction;
    };
    struct ConvexResultCallback {
        btScalar m_closestHitFraction = btScalar(1.);
        int m_collisionFilterGroup = btBroadphaseProxy::DefaultFilter, m_collisionFilterMask = btBroadphaseProxy::AllFilter;
        virtual ~ConvexResultCallback() = default;
        bool hasHit() const { return m_closestHitFraction < btScalar(1.); }
        virtual bool needsCollision(btBroadphaseProxy *proxy0) const {
            bool collides = (proxy0->m_collisionFilterGroup & m_collisionFilterMask) != 0;
            collides = collides && (m_collisionFilterGroup & proxy0->m_collisionFilterMask);
            return collides;
        }
        virtual btScalar addSingleResult(LocalConvexResult &convexResult, bool normalInWorldSpace) = 0;
    };
    struct ClosestConvexResultCallback : public ConvexResultCallback {
        ClosestConvexResultCallback(const btVector3 &from, const btVector3 &to) : m_convexFromWorld(from), m_convexToWorld(to), m_hitCollisionObject(nullptr) {}
        btVector3 m_convexFromWorld, m_convexToWorld, m_hitNormalWorld, m_hitPointWorld;
        const btCollisionObject *m_hitCollisionObject;
        btScalar addSingleResult(LocalConvexResult &convexResult, bool normalInWorldSpace) override {
            m_closestHitFraction = convexResult.m_hitFraction;
            m_hitCollisionObject = convexResult.m_hitCollisionObject;
            if (normalInWorldSpace) m_hitNormalWorld = convexResult.m_hitNormalLocal;
            else m_hitNormalWorld = m_hitCollisionObject->getWorldTransform().getBasis() * convexResult.m_hitNormalLocal;
            m_hitPointWorld = convexResult.m_hitPointLocal;
            return convexResult.m_hitFraction;
        }
    };
    struct LocalRayResult {
        const btCollisionObject *m_collisionObject;
        LocalShapeInfo *m_localShapeInfo;
        btVector3 m_hitNormalLocal;
        btScalar m_hitFraction;
    };
    struct RayResultCallback {
        btScalar m_closestHitFraction = btScalar(1.);
        const btCollisionObject *m_collisionObject = nullptr;
        virtual ~RayResultCallback() = default;
        virtual btScalar addSingleResult(LocalRayResult &rayResult, bool normalInWorldSpace) = 0;
    };
    struct ClosestRayResultCallback : public RayResultCallback {
        ClosestRayResultCallback(const btVector3 &from, const btVector3 &to) : m_rayFromWorld(from), m_rayToWorld(to) {}
        btVector3 m_rayFromWorld, m_rayToWorld, m_hitNormalWorld, m_hitPointWorld;
        btScalar addSingleResult(LocalRayResult &rayResult, bool) override {
            m_closestHitFraction = rayResult.m_hitFraction;
            m_collisionObject = rayResult.m_collisionObject;
            m_hitNormalWorld = rayResult.m_hitNormalLocal;
            return rayResult.m_hitFraction;
        }
    };

    virtual ~btCollisionWorld() = default;
    btDispatcherInfo &getDispatchInfo() { return m_dispatchInfo; }
    btDispatcher *getDispatcher() { return m_dispatcher; }
    btBroadphaseInterface *getBroadphase() { return m_broadphase; }

    // the stand-in's visiting order: static bodies in joining order, then character objects in joining order
    std::vector<btCollisionObject *> ordered() const {
        std::vector<btCollisionObject *> o;
        for (auto *c : m_objects) if (!(c->getCollisionFlags() & btCollisionObject::CF_CHARACTER_OBJECT)) o.push_back(c);
        for (auto *c : m_objects) if (c->getCollisionFlags() & btCollisionObject::CF_CHARACTER_OBJECT) o.push_back(c);
        return o;
    }
    // one capsule sweep against everything, answered by the oracle's analytic narrow phase.  The loop is btCollisionWorld's
    // (objectQuerySingle: filter through the callback, report only fractions below the callback's current closest one).
    void sweepCapsule(const btCollisionObject *me, const btVector3 &from, const btVector3 &to, ConvexResultCallback &cb) const {
        const orc::Vec3 f = toOrc(from), t = toOrc(to), d = t - f;
        for (btCollisionObject *o : ordered()) {
            if (o == me) continue;  // (the ghost never overlaps itself; the reference's callback would reject it as well)
            if (!cb.needsCollision(o->getBroadphaseHandle())) continue;
            const orc::Collider c = o->collider();
            if (orc::sweepBroadphaseMiss(c, f, t)) continue;
            float frac;
            orc::Vec3 n;
            if (!orc::sweepNarrow(c, f, d, frac, n)) continue;
            if (!(frac < cb.m_closestHitFraction)) continue;
            LocalConvexResult r(o, nullptr, fromOrc(n), btVector3(0, 0, 0), frac);
            cb.addSingleResult(r, true);
        }
    }
    void convexSweepTest(const btConvexShape *, const btTransform &from, const btTransform &to, ConvexResultCallback &cb, btScalar = 0) const {
        sweepCapsule(nullptr, from.getOrigin(), to.getOrigin(), cb);
    }

    std::vector<btCollisionObject *> m_objects;
    btDispatcherInfo m_dispatchInfo;
    btDispatcher *m_dispatcher = nullptr;
    btBroadphaseInterface *m_broadphase = nullptr;
};

class btGhostObject : public btCollisionObject {
public:
    void convexSweepTest(const btConvexShape *shape, const btTransform &from, const btTransform &to, btCollisionWorld::ConvexResultCallback &cb, btScalar allowed = 0) const {
        // the analytic narrow phase is defined for the agents' capsule and the world's default allowed penetration
        if (shape->kind != btCollisionShape::CAPSULE || shape->m_dims.x() != orc::kCapsuleRadius || shape->m_dims.y() != 2 * orc::kCapsuleHalfHeight ||
            allowed != orc::kAllowedCcdPenetration) {
            std::fprintf(stderr, "mini_bullet: sweep with a shape / tolerance the oracle's narrow phase is not defined for\n");
            std::abort();
        }
        m_world->sweepCapsule(this, from.getOrigin(), to.getOrigin(), cb);
    }
};
class btPairCachingGhostObject : public btGhostObject {
public:
    btPairCachingGhostObject() { m_cache.m_owner = this; }
    btHashedOverlappingPairCache *getOverlappingPairCache() { return &m_cache; }
    btHashedOverlappingPairCache m_cache;
};

inline void btDispatcher::dispatchAllCollisionPairs(btOverlappingPairCache *cache, const btDispatcherInfo &, btDispatcher *) {
    auto *hc = dynamic_cast<btHashedOverlappingPairCache *>(cache);
    if (!hc || !hc->m_owner) return;
    btPairCachingGhostObject *ghost = hc->m_owner;
    hc->m_pairs.clear();
    hc->m_algorithms.clear();
    const orc::Vec3 p = toOrc(ghost->getWorldTransform().getOrigin());
    for (btCollisionObject *o : ghost->m_world->ordered()) {
        if (o == ghost) continue;
        orc::Vec3 n;
        const float dist = orc::capsuleDistance(o->collider(), p, n);
        auto algo = std::make_unique<btCollisionAlgorithm>();
        algo->m_manifold.m_body0 = ghost;  // normal on B (the other object) points at the ghost: directionSign = -1 in the controller
        algo->m_manifold.m_body1 = o;
        btManifoldPoint pt;
        pt.m_normalWorldOnB = fromOrc(n);
        pt.m_distance1 = dist;
        algo->m_manifold.m_points.push_back(pt);
        btBroadphasePair pair;
        pair.m_pProxy0 = ghost->getBroadphaseHandle();
        pair.m_pProxy1 = o->getBroadphaseHandle();
        pair.m_algorithm = algo.get();
        hc->m_pairs.push_back(pair);
        hc->m_algorithms.push_back(std::move(algo));
    }
}

class btDynamicsWorld : public btCollisionWorld {
public:
    void addRigidBody(btRigidBody *body) {  // btDiscreteDynamicsWorld::addRigidBody: static bodies get StaticFilter / everything but static
        const bool isDynamic = !(body->getCollisionFlags() & (btCollisionObject::CF_STATIC_OBJECT | btCollisionObject::CF_KINEMATIC_OBJECT));
        addCollisionObject(body, isDynamic ? int(btBroadphaseProxy::DefaultFilter) : int(btBroadphaseProxy::StaticFilter),
                           isDynamic ? int(btBroadphaseProxy::AllFilter) : int(btBroadphaseProxy::AllFilter ^ btBroadphaseProxy::StaticFilter));
    }
    void removeRigidBody(btRigidBody *body) { removeCollisionObject(body); }
    void addCollisionObject(btCollisionObject *o, int group = btBroadphaseProxy::DefaultFilter, int mask = btBroadphaseProxy::AllFilter) {
        o->m_world = this;
        o->getBroadphaseHandle()->m_collisionFilterGroup = group;
        o->getBroadphaseHandle()->m_collisionFilterMask = mask;
        m_objects.push_back(o);
    }
    void removeCollisionObject(btCollisionObject *o) { m_objects.erase(std::remove(m_objects.begin(), m_objects.end(), o), m_objects.end()); }
    void addAction(btActionInterface *a) { m_actions.push_back(a); }
    void removeAction(btActionInterface *a) { m_actions.erase(std::remove(m_actions.begin(), m_actions.end(), a), m_actions.end()); }
    // btDiscreteDynamicsWorld::stepSimulation with maxSubSteps > 0: accumulate, run whole fixed steps (clamped), actions in joining order
    int stepSimulation(btScalar timeStep, int maxSubSteps = 1, btScalar fixedTimeStep = btScalar(1.) / btScalar(60.)) {
        int numSimulationSubSteps = 0;
        m_localTime += timeStep;
        if (m_localTime >= fixedTimeStep) {
            numSimulationSubSteps = int(m_localTime / fixedTimeStep);
            m_localTime -= numSimulationSubSteps * fixedTimeStep;
        }
        if (numSimulationSubSteps) {
            const int clamped = numSimulationSubSteps > maxSubSteps ? maxSubSteps : numSimulationSubSteps;
            for (int i = 0; i < clamped; ++i)
                for (size_t a = 0; a < m_actions.size(); ++a) m_actions[a]->updateAction(this, fixedTimeStep);
        }
        return numSimulationSubSteps;
    }
    std::vector<btActionInterface *> m_actions;
    btScalar m_localTime = 0;
};
class btDiscreteDynamicsWorld : public btDynamicsWorld {
public:
    btDiscreteDynamicsWorld(btDispatcher *dispatcher, btBroadphaseInterface *pairCache, btConstraintSolver *, btCollisionConfiguration *) {
        m_dispatcher = dispatcher;
        m_broadphase = pairCache;
    }
};
