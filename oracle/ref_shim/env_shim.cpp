// ORACLE tooling -- NOT product code.  Runs the reference's REAL env library -- env/src/env.cpp, agent.cpp,
// kinematic_character_controller.cpp, env/physics.hpp's RigidBody, magnum-integration's MotionState, and every scenario source --
// compiled unmodified, in place, on the Bullet stand-in of ref_shim/mini_bullet (LinearMath + world containers restated; the narrow
// phase answered by the oracle's analytic definitions).  Whole episodes run here exactly as Env::reset() / Env::step() drive them;
// tests/test_ref_shim.py compares every tick with the oracle: rewards, timers, done, true objectives, every drawable's absolute
// matrix (the agents' bodies, eyes and HUD bars carry their position, heading and camera pitch) and every collider.
//
// One substitution, as in scen_shim.cpp: the maze library's SpanningtreeAlgorithm constructor takes its seed from the test.
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include <mazes/spanningtreealgorithm.h>

#include <Magnum/SceneGraph/Camera.h>

#include <env/env.hpp>
#include <env/env_renderer.hpp>
#include <env/scenario.hpp>
#include <env/vector_env.hpp>
#include <scenarios/scenario_collect.hpp>
#include <scenarios/scenario_empty.hpp>
#include <scenarios/scenario_hex_explore.hpp>
#include <scenarios/scenario_hex_memory.hpp>
#include <scenarios/scenario_obstacles.hpp>
#include <scenarios/scenario_rearrange.hpp>
#include <scenarios/scenario_sokoban.hpp>
#include <scenarios/scenario_tower_building.hpp>

// Second substitution: inside this library libm's sinf / cosf / sincosf (reached from Magnum's Matrix4::rotation*, i.e. the agents'
// display transforms) are the correctly rounded float(fn(double)) the oracle and the device define them as (orc_math.hpp:26-29).
// glibc's own float versions differ from that by one ulp for roughly one argument in a few thousand, which shows up as a last-bit
// difference of a heading matrix about once per hundred ticks.  (The library is linked -Bsymbolic-functions so that these win.)
extern "C" {
float sinf(float x) noexcept { return float(std::sin(double(x))); }
float cosf(float x) noexcept { return float(std::cos(double(x))); }
void sincosf(float x, float *s, float *c) noexcept { *s = float(std::sin(double(x))); *c = float(std::cos(double(x))); }
}

static unsigned g_mazeSeed = 0;
static std::deque<unsigned> g_mazeSeedQueue;  // VectorEnv runs: one entry per env about to be reset, in reset order
SpanningtreeAlgorithm::SpanningtreeAlgorithm() {
    unsigned seed = g_mazeSeed;
    if (!g_mazeSeedQueue.empty()) { seed = g_mazeSeedQueue.front(); g_mazeSeedQueue.pop_front(); }
    generator = std::mt19937(seed);
}

using namespace Megaverse;

namespace {

template <typename T> void reg(const std::string &name) { Scenario::registerScenario(name, Scenario::scenarioFactory<T>); }

void registerScenarios() {  // scenarios/init.hpp:28-56 without the experimental Football / BoxAGone (dynamic Bullet bodies)
    static bool done = false;
    if (done) return;
    done = true;
    setLogLevel(WARNING);
    reg<TowerBuildingScenario>("TowerBuilding");
    reg<ObstaclesEasyScenario>("ObstaclesEasy");
    reg<ObstaclesMediumScenario>("ObstaclesMedium");
    reg<ObstaclesHardScenario>("ObstaclesHard");
    reg<ObstaclesOnlyWallsScenario>("ObstaclesWalls");
    reg<ObstaclesOnlyStepsScenario>("ObstaclesSteps");
    reg<ObstaclesOnlyLavaScenario>("ObstaclesLava");
    reg<CollectScenario>("Collect");
    reg<SokobanScenario>("Sokoban");
    reg<RearrangeScenario>("Rearrange");
    reg<HexExploreScenario>("HexExplore");
    reg<HexMemoryScenario>("HexMemory");
    reg<EmptyScenario>("Empty");
}

uint32_t bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

// read access to two private members without touching the reference's headers: explicit instantiation may name private members
template <typename Tag, typename Tag::type M> struct Rob { friend typename Tag::type get(Tag) { return M; } };
struct EnvStateTag { typedef Env::EnvState Env::*type; friend type get(EnvStateTag); };
template struct Rob<EnvStateTag, &Env::state>;
struct GhostTag { typedef btPairCachingGhostObject DefaultKinematicAgent::*type; friend type get(GhostTag); };
template struct Rob<GhostTag, &DefaultKinematicAgent::ghostObject>;

struct Handle {
    Handle(const std::string &name, int numAgents, const FloatParams &fp) : env{name, numAgents, fp} {}
    Env env;
    Env::EnvState &state() { return env.*get(EnvStateTag()); }
};

}  // namespace

extern "C" {

void *ref_env_create(const char *name, int numAgents, const char **keys, const float *vals, int nparams) {
    registerScenarios();
    FloatParams fp;
    for (int i = 0; i < nparams; ++i) fp[keys[i]] = vals[i];
    return new Handle(name, numAgents, fp);
}
void ref_env_destroy(void *p) { delete static_cast<Handle *>(p); }
void ref_env_seed(void *p, int seedValue) { static_cast<Handle *>(p)->env.seed(seedValue); }
void ref_env_reset(void *p, unsigned mazeSeedXor) {
    auto &h = *static_cast<Handle *>(p);
    // Env::reset() re-seeds its rng from one draw (env.cpp:62-63); the maze seed the oracle derives from that value has to be known
    // BEFORE reset() runs the scenario, so the draw is replayed on a copy of the generator
    Rng copy = h.env.getRng();
    const auto seed = randRange(0, 1 << 30, copy);
    g_mazeSeed = unsigned(seed) ^ mazeSeedXor;
    h.env.reset();
}
void ref_env_step(void *p, const int *actions) {
    auto &h = *static_cast<Handle *>(p);
    for (int i = 0; i < h.env.getNumAgents(); ++i) h.env.setAction(i, Action(actions[i]));
    h.env.step();
}
// test hook matching orc_scen_warp: KinematicCharacterController::warp through the agent's own teleport() (rotation back to
// identity), then a heading the way DefaultKinematicAgent's constructor sets one (agent.cpp:40-44)
void ref_env_warp(void *p, int agent, float x, float y, float z, float yaw) {
    auto &h = *static_cast<Handle *>(p);
    auto *a = dynamic_cast<DefaultKinematicAgent *>(h.env.getAgents()[size_t(agent)]);
    a->teleport(btVector3(x, y, z));
    btPairCachingGhostObject &ghost = a->*get(GhostTag());
    btTransform t = ghost.getWorldTransform();
    t.setRotation(btQuaternion(btVector3(0, 1, 0), yaw));
    ghost.setWorldTransform(t);
}

// Env::actionSpaceSizes (env.cpp:33), the table MegaverseGym::setActions (megaverse.cpp:100-116) walks
int ref_action_space_sizes(int *out, int cap) {
    const auto &v = Env::actionSpaceSizes;
    if (int(v.size()) > cap) return -int(v.size());
    std::copy(v.begin(), v.end(), out);
    return int(v.size());
}

// a scenario's reward shaping and float parameters right after construction, same text layout as mv_debug_defaults
int ref_env_defaults(void *p, char *out, int cap) {
    auto &h = *static_cast<Handle *>(p);
    std::string text;
    char line[160];
    auto put = [&](char tag, const std::string &k, float v) {
        std::snprintf(line, sizeof(line), "%c %s=%08x\n", tag, k.c_str(), bits(v));
        text += line;
    };
    for (auto &kv : h.env.getScenario().getRewardShaping(0)) put('R', kv.first, kv.second);
    for (auto &kv : h.env.getScenario().getFloatParams()) put('P', kv.first, kv.second);
    if (int(text.size()) + 1 > cap) return -int(text.size()) - 1;
    std::memcpy(out, text.c_str(), text.size() + 1);
    return int(text.size());
}

// what the renderer reads per agent (v4r_env_renderer.cpp:303-314): Camera3D::cameraMatrix(), 16 floats column-major
void ref_env_views(void *p, float *out) {
    auto &h = *static_cast<Handle *>(p);
    for (int i = 0; i < h.env.getNumAgents(); ++i) {
        const auto m = h.env.getAgents()[size_t(i)]->getCamera()->cameraMatrix();
        std::memcpy(out + 16 * i, m.data(), 64);
    }
}

// Same layout as orc_scenario_dump (oracle/orc_api.cpp).  Floats as bit patterns.
static int dumpEnv(Env &env, uint32_t *out, int cap) {
    std::vector<uint32_t> o;
    o.push_back(bits(env.episodeLengthSec()));
    o.push_back(env.isDone() ? 1u : 0u);
    o.push_back(bits((env.*get(EnvStateTag())).currEpisodeSec));
    o.push_back(uint32_t(env.getNumAgents()));
    for (int i = 0; i < env.getNumAgents(); ++i) {
        o.push_back(bits(env.getLastReward(i)));
        o.push_back(bits(env.getTotalReward(i)));
        o.push_back(bits(env.trueObjective(i)));
    }
    const auto &drawables = env.getDrawables();
    size_t n = 0;
    for (auto &[type, list] : drawables) n += list.size();
    o.push_back(uint32_t(n));
    for (auto &[type, list] : drawables)
        for (auto &d : list) {
            o.push_back(uint32_t(type));
            const auto c = d.color;
            o.push_back((uint32_t(c.r() * 255.0f + 0.5f) << 16) | (uint32_t(c.g() * 255.0f + 0.5f) << 8) | uint32_t(c.b() * 255.0f + 0.5f));
            const auto m = d.objectPtr->absoluteTransformationMatrix();
            for (int k = 0; k < 16; ++k) o.push_back(bits(m.data()[k]));
        }
    std::vector<btCollisionObject *> statics;
    for (auto *c : env.getPhysics().bWorld.m_objects)
        if (!(c->getCollisionFlags() & btCollisionObject::CF_CHARACTER_OBJECT)) statics.push_back(c);
    o.push_back(uint32_t(statics.size()));
    for (auto *c : statics) {
        const auto col = c->collider();
        o.push_back(bits(col.c.x)), o.push_back(bits(col.c.y)), o.push_back(bits(col.c.z));
        o.push_back(bits(col.h.x)), o.push_back(bits(col.h.y)), o.push_back(bits(col.h.z));
        const auto &b = c->getWorldTransform().getBasis();
        o.push_back(bits(b[0].x())), o.push_back(bits(b[2].x()));
        o.push_back(col.enabled ? 1u : 0u);
    }
    if (int(o.size()) > cap) return -int(o.size());
    std::copy(o.begin(), o.end(), out);
    return int(o.size());
}
int ref_env_dump(void *p, uint32_t *out, int cap) { return dumpEnv(static_cast<Handle *>(p)->env, out, cap); }

// ---- the reference's VectorEnv (vector_env.cpp, compiled as is) over several envs, read the way MegaverseGym reads it
// (megaverse.cpp:118-143): step(), then getLastReward per agent (zeroed by the reset of a finished env), done, trueObjectives
namespace {
unsigned predictedMazeSeed(Env &env, unsigned mazeSeedXor) {
    Rng copy = env.getRng();
    return unsigned(randRange(0, 1 << 30, copy)) ^ mazeSeedXor;
}
struct NullRenderer : EnvRenderer {
    unsigned mazeSeedXor = 0;
    void reset(Env &, int) override {}
    // runs right after each env's step (vector_env.cpp:49-52): a finished env is about to be reset by VectorEnv::step's loop
    void preDraw(Env &env, int) override { if (env.isDone()) g_mazeSeedQueue.push_back(predictedMazeSeed(env, mazeSeedXor)); }
    void draw(Envs &) override {}
    const uint8_t *getObservation(int, int) const override { return nullptr; }
    Overview *getOverview() override { return nullptr; }
};
struct VecHandle {
    Envs envs;
    NullRenderer renderer;
    std::unique_ptr<VectorEnv> vec;
};
}  // namespace

void *ref_vec_create(const char *name, int numEnvs, int numAgents, const char **keys, const float *vals, int nparams, unsigned mazeSeedXor) {
    registerScenarios();
    FloatParams fp;
    for (int i = 0; i < nparams; ++i) fp[keys[i]] = vals[i];
    auto v = std::make_unique<VecHandle>();
    for (int i = 0; i < numEnvs; ++i) v->envs.emplace_back(std::make_unique<Env>(name, numAgents, fp));
    v->renderer.mazeSeedXor = mazeSeedXor;
    v->vec = std::make_unique<VectorEnv>(v->envs, v->renderer, 1);
    return v.release();
}
void ref_vec_destroy(void *p) {
    auto *v = static_cast<VecHandle *>(p);
    v->vec->close();
    delete v;
}
void ref_vec_seed_env(void *p, int env, int seedValue) { static_cast<VecHandle *>(p)->envs[size_t(env)]->seed(seedValue); }
void ref_vec_reset(void *p) {
    auto *v = static_cast<VecHandle *>(p);
    g_mazeSeedQueue.clear();
    for (auto &e : v->envs) g_mazeSeedQueue.push_back(predictedMazeSeed(*e, v->renderer.mazeSeedXor));  // VectorEnv::reset resets in index order
    v->vec->reset();
    g_mazeSeedQueue.clear();  // (scenarios without a maze never consume their entries)
}
// masks[env * A + agent]; out: rewards[N], dones[E], trueObjectives[N] (valid where done)
void ref_vec_step(void *p, const int *masks, float *rewards, uint8_t *dones, float *trueObjectives) {
    auto *v = static_cast<VecHandle *>(p);
    const int E = int(v->envs.size()), A = v->envs[0]->getNumAgents();
    for (int e = 0; e < E; ++e)
        for (int a = 0; a < A; ++a) v->envs[size_t(e)]->setAction(a, Action(masks[e * A + a]));
    g_mazeSeedQueue.clear();
    v->vec->step();
    g_mazeSeedQueue.clear();
    for (int e = 0; e < E; ++e) {
        dones[e] = v->vec->done[size_t(e)] ? 1 : 0;
        for (int a = 0; a < A; ++a) {
            rewards[e * A + a] = v->envs[size_t(e)]->getLastReward(a);
            trueObjectives[e * A + a] = v->vec->trueObjectives[size_t(e)][size_t(a)];
        }
    }
}
int ref_vec_dump(void *p, int env, uint32_t *out, int cap) { return dumpEnv(*static_cast<VecHandle *>(p)->envs[size_t(env)], out, cap); }

}  // extern "C"
