// ORACLE tooling.  Stand-in for the reference's env/env.hpp when compiling scenarios/platforms.hpp, component_voxel_grid.hpp and
// component_object_stacking.hpp alone.  The real header pulls in Bullet (physics.hpp), which this image does not have.  What the
// three scenario headers need from it is only a vocabulary: the FloatParams alias (env.hpp:85), the Action bit for Interact
// (env.hpp:22-42), and the NAMES Env / Env::EnvState / AbstractAgent / RigidBody with the members those headers touch.  The bodies
// below do no physics: RigidBody just remembers whether it collides -- the logic under test is the reference's own header.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <Magnum/SceneGraph/Object.h>
#include <Magnum/SceneGraph/Scene.h>
#include <util/magnum.hpp>
#include <env/const.hpp>
#include <env/voxel_state.hpp>

// names of Bullet types that appear in non-dependent code of the scenario headers (never instantiated by the shim)
struct btVector3 { float x, y, z; };
struct btCollisionShape { virtual ~btCollisionShape() = default; };
struct btBoxShape : btCollisionShape { explicit btBoxShape(btVector3) {} };
struct btDynamicsWorld {};

namespace Megaverse {

using FloatParams = std::map<std::string, float>;

enum class Action { Idle = 0, Interact = 1 << 8 };  // env.hpp:22-42 (only the bit the stacking component tests)
inline Action operator&(Action a, Action b) { return Action(int(a) & int(b)); }
inline bool operator!(Action a) { return int(a) == 0; }

enum class DrawableType { First = 0, Box = 0, Capsule = 1, Sphere = 2, Cone = 3, Cylinder = 4, NumTypes };  // env.hpp:57-67
struct SceneObjectInfo {  // env.hpp:70-80
    SceneObjectInfo(Object3D *objectPtr, const Magnum::Color3 &color) : objectPtr{objectPtr}, color{color} {}
    Object3D *objectPtr;
    Magnum::Color3 color;
};
using DrawablesMap = std::map<DrawableType, std::vector<SceneObjectInfo>>;

class RigidBody : public Object3D {  // physics.hpp:19-102 without Bullet
public:
    RigidBody(Object3D *parent, Magnum::Float, btCollisionShape *, btDynamicsWorld &) : Object3D{parent} {}
    void setCollisionScale(const Magnum::Vector3 &s) { collisionScale = s; }
    void setCollisionOffset(const Magnum::Vector3 &o) { collisionOffset = o; }
    // physics.hpp:69-74: world transform = (rotation, translation + offset), local scaling = scaling * collisionScale -- kept as numbers
    void syncPose() {
        const auto &m = absoluteTransformationMatrix();
        colliderOrigin = m.translation() + collisionOffset;
        colliderScaling = m.scaling() * collisionScale;
        ++syncs;
    }
    Magnum::Vector3 collisionScale{1, 1, 1}, collisionOffset{0, 0, 0}, colliderOrigin{0, 0, 0}, colliderScaling{0, 0, 0};
    void toggleCollision() { collides = !collides; }
    bool colliding() const { return collides; }
    bool collides = true;
    int syncs = 0;
};

class AbstractAgent : public Object3D {  // agent.hpp:27-60 (the members the scenario components use)
public:
    explicit AbstractAgent(Object3D *parent) : Object3D{parent} {}
    virtual Object3D *interactLocation() = 0;
};

class Env {
public:
    struct Physics { btDynamicsWorld bWorld; std::vector<std::unique_ptr<btCollisionShape>> collisionShapes; };
    struct EnvState {
        std::vector<Action> currAction;
        std::vector<AbstractAgent *> agents;
        std::unique_ptr<Scene3D> scene;
        std::unique_ptr<Physics> physics;
    };
    int getNumAgents() const { return numAgents; }
    int numAgents = 0;
};

}  // namespace Megaverse
