// ORACLE tooling.  Stand-in for the reference's env/env.hpp + env/agent.hpp + env/physics.hpp, used to compile the reference's REAL
// scenario sources (scenario_*.cpp, component_*.hpp/.cpp, layout_utils.cpp, env/scenario.hpp, scenario_default.hpp) in place
// without Bullet.  It supplies only the VOCABULARY those sources use -- every rule under test is the reference's own code:
//   Action bits (env.hpp:22-42), DrawableType / SceneObjectInfo / DrawablesMap (env.hpp:57-84), FloatParams (env.hpp:85),
//   Env::EnvState's fields (env.hpp:112-158) and Env's accessors (env.hpp:166-228),
//   RigidBody (physics.hpp:19-102): keeps the numbers syncPose() would hand to Bullet instead of handing them over,
//   AbstractAgent / DefaultKinematicAgent (agent.hpp:20-130): a POSED puppet -- the test sets its transform every tick from the
//   oracle's kinematics, so what is compared is the scenario logic given identical agent poses.
#pragma once
#include <algorithm>
#include <cfloat>   // the Bullet headers the real env.hpp includes bring these two in
#include <cstring>
#include <map>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include <Magnum/SceneGraph/Object.h>
#include <Magnum/SceneGraph/Scene.h>
#include <Magnum/SceneGraph/SceneGraph.h>
#include <util/magnum.hpp>
#include <util/util.hpp>
#include <env/const.hpp>
#include <env/voxel_state.hpp>

// names of Bullet types that appear in the scenario sources (shape objects are created and parked, never used)
struct btVector3 {
    btVector3() = default;
    btVector3(float x, float y, float z) : v{x, y, z} {}
    float x() const { return v[0]; }
    float y() const { return v[1]; }
    float z() const { return v[2]; }
    float v[3] = {0, 0, 0};
};
using btScalar = float;
struct btCollisionShape { virtual ~btCollisionShape() = default; };
struct btBoxShape : btCollisionShape { explicit btBoxShape(btVector3) {} };
struct btDynamicsWorld {};

namespace Megaverse {

class Scenario;
using FloatParams = std::map<std::string, float>;

enum class Action {  // env.hpp:22-42
    Idle = 0, Left = 1 << 1, Right = 1 << 2, Forward = 1 << 3, Backward = 1 << 4, LookLeft = 1 << 5, LookRight = 1 << 6,
    Jump = 1 << 7, Interact = 1 << 8, LookDown = 1 << 9, LookUp = 1 << 10, NumActions = 11,
};
inline Action operator|(Action a, Action b) { return Action(int(a) | int(b)); }
inline Action operator&(Action a, Action b) { return Action(int(a) & int(b)); }
inline Action operator~(Action a) { return Action(~int(a)); }
inline bool operator!(Action a) { return a == Action::Idle; }

enum class DrawableType { First = 0, Box = 0, Capsule = 1, Sphere = 2, Cone = 3, Cylinder = 4, NumTypes };  // env.hpp:57-67
struct SceneObjectInfo {  // env.hpp:70-80
    SceneObjectInfo(Object3D *objectPtr, const Magnum::Color3 &color) : objectPtr{objectPtr}, color{color} {}
    Object3D *objectPtr;
    Magnum::Color3 color;
};
using DrawablesMap = std::map<DrawableType, std::vector<SceneObjectInfo>>;
using RewardShaping = std::map<std::string, float>;

class RigidBody;
// creation-ordered registry of the live bodies of the current episode (the order Bullet's world would hold them in)
inline std::vector<RigidBody *> &standinBodies() { static std::vector<RigidBody *> v; return v; }

class RigidBody : public Object3D {  // physics.hpp:19-102 without Bullet
public:
    RigidBody(Object3D *parent, Magnum::Float, btCollisionShape *, btDynamicsWorld &) : Object3D{parent} { standinBodies().push_back(this); }
    ~RigidBody() override {
        auto &v = standinBodies();
        v.erase(std::remove(v.begin(), v.end(), this), v.end());
    }
    bool colliding() const { return collides; }
    void setCollisionScale(const Magnum::Vector3 &s) { collisionScale = s; }
    void setCollisionOffset(const Magnum::Vector3 &o) { collisionOffset = o; }
    // physics.hpp:69-74: world transform = (rotation, translation + offset), local scaling = scaling * collisionScale -- kept as numbers
    void syncPose() {
        const auto &m = absoluteTransformationMatrix();
        colliderOrigin = m.translation() + collisionOffset;
        colliderScaling = m.scaling() * collisionScale;
        const auto rs = m.rotationScaling();
        colliderAxisX = rs[0].normalized();
        colliderAxisZ = rs[2].normalized();
        ++syncs;
    }
    void toggleCollision() { collides = !collides; }
    Magnum::Vector3 collisionScale{1, 1, 1}, collisionOffset{0, 0, 0}, colliderOrigin{0, 0, 0}, colliderScaling{0, 0, 0};
    Magnum::Vector3 colliderAxisX{1, 0, 0}, colliderAxisZ{0, 0, 1};
    bool collides = true;
    int syncs = 0;
};

class AbstractAgent : public Object3D {  // agent.hpp:20-62
public:
    explicit AbstractAgent(Object3D *parent, btDynamicsWorld &, float verticalLookLimitRad)
    : Object3D{parent}, verticalLookLimitRad{verticalLookLimitRad} {}
    virtual void updateTransform() = 0;
    virtual bool onGround() const = 0;
    virtual void teleport(const btVector3 &position) = 0;
    virtual float getAgentHeight() = 0;
    virtual Object3D *getCameraObject() = 0;
    virtual Object3D *interactLocation() = 0;
protected:
    float verticalLookLimitRad = 0.0f;
};

class DefaultKinematicAgent : public AbstractAgent {  // agent.hpp:65-130, agent.cpp:24-63 -- scene-graph part only
public:
    explicit DefaultKinematicAgent(Object3D *parent, btDynamicsWorld &bWorld, const Magnum::Vector3 &startingPosition,
                                   float rotationRad, float verticalLookLimitRad)
    : AbstractAgent{parent, bWorld, verticalLookLimitRad}
    , cameraObject{&(addChild<Object3D>())}
    , pickupSpot{&(cameraObject->addChild<Object3D>())}
    , startingPosition{startingPosition}, rotationRad{rotationRad} {
        cameraObject->translate(Magnum::Vector3{0, 0.41f, 0});  // agent.cpp:31
        pickupSpot->translate({0.0f, -0.44f, -1.0f});           // agent.cpp:38
    }
    void updateTransform() override {}  // the pose is set by the test
    bool onGround() const override { return grounded; }
    void teleport(const btVector3 &p) override { teleports.push_back({p.x(), p.y(), p.z()}); }
    float getAgentHeight() override { return 1.75f; }  // agent.hpp:110
    Object3D *getCameraObject() override { return cameraObject; }
    Object3D *interactLocation() override { return pickupSpot; }

    Object3D *cameraObject, *pickupSpot;
    Magnum::Vector3 startingPosition;
    float rotationRad;
    bool grounded = true;
    std::vector<Magnum::Vector3> teleports;
};

using Agents = std::vector<AbstractAgent *>;

class Env {
public:
    struct EnvPhysics { btDynamicsWorld bWorld; std::vector<std::unique_ptr<btCollisionShape>> collisionShapes; };
    struct EnvState {  // env.hpp:112-158
        explicit EnvState(int numAgents)
        : physics{std::make_unique<EnvPhysics>()}, currAction(size_t(numAgents), Action::Idle), lastReward(size_t(numAgents), 0)
        , totalReward(size_t(numAgents), 0.0f) {}
        void reset() {  // env.hpp:123-139
            done = false, currEpisodeSec = 0, numFrames = 0;
            std::fill(currAction.begin(), currAction.end(), Action::Idle);
            std::fill(lastReward.begin(), lastReward.end(), 0.0f);
            std::fill(totalReward.begin(), totalReward.end(), 0.0f);
            scene = std::make_unique<Scene3D>();
            agents.clear();
            physics = std::make_unique<EnvPhysics>();
        }
        std::unique_ptr<EnvPhysics> physics;
        bool done = false;
        int numFrames = 0;
        float currEpisodeSec = 0;
        float simulationStepSeconds = 1.0f / 15.0f;
        float lastFrameDurationSec = simulationStepSeconds;
        std::vector<Action> currAction;
        std::vector<float> lastReward, totalReward;
        std::unique_ptr<Scene3D> scene;
        Agents agents;
        Rng rng{0};
    };

    explicit Env(int numAgents) : state{numAgents}, numAgents{numAgents} {}
    int getNumAgents() const { return numAgents; }
    Agents &getAgents() { return state.agents; }
    float episodeLengthSec() const;  // defined by the shim once Scenario is complete (env.cpp:163-166)
    float remainingTimeFraction() const {  // env.hpp:224-228
        const auto len = episodeLengthSec();
        return std::max(0.0f, (len - state.currEpisodeSec) / len);
    }

    EnvState state;
    int numAgents;
    Scenario *scenarioPtr = nullptr;
};

}  // namespace Megaverse
