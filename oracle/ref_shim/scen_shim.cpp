// ORACLE tooling -- NOT product code.  Drives the reference's REAL scenario sources (scenarios/src/scenario_*.cpp,
// component_hexagonal_maze.cpp, layout_utils.cpp, the component headers, env/scenario.hpp, scenarios/scenario_default.hpp), compiled in
// place from /root/reference against the Bullet-free stand-ins in ref_shim/inc_scen/env/, through the same sequence Env::reset()
// (env.cpp:58-77) and Env::step() (env.cpp:84-152) run them in.  The agents are posed puppets: every tick the test copies the
// oracle's agent transforms in, so the comparison isolates the scenario rules -- level construction, drawables, colliders handed to
// Bullet, rewards, timers, true objectives -- given identical agent poses.  tests/test_ref_shim.py compares the dumps with orc_scenario_dump.
//
// One substitution: the maze library's SpanningtreeAlgorithm constructor (mazes/src/spanningtreealgorithm.cpp:3-5) seeds its mt19937
// from std::random_device; the definition below seeds it from a value the test sets, so the hexagonal scenarios are repeatable.
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include <mazes/spanningtreealgorithm.h>

#include <env/scenario.hpp>
#include <scenarios/scenario_collect.hpp>
#include <scenarios/scenario_hex_explore.hpp>
#include <scenarios/scenario_hex_memory.hpp>
#include <scenarios/scenario_obstacles.hpp>
#include <scenarios/scenario_rearrange.hpp>
#include <scenarios/scenario_sokoban.hpp>
#include <scenarios/scenario_tower_building.hpp>

static unsigned g_mazeSeed = 0;
SpanningtreeAlgorithm::SpanningtreeAlgorithm() { generator = std::mt19937(g_mazeSeed); }

using namespace Megaverse;

float Env::episodeLengthSec() const { return scenarioPtr->episodeLengthSec(); }  // env.cpp:163-166

namespace {

struct Handle {
    explicit Handle(int numAgents) : env{numAgents} {}
    Env env;
    std::unique_ptr<Scenario> scenario;
    DrawablesMap drawables;
};

template <typename T> std::unique_ptr<Scenario> make(const std::string &name, Env &env) { return std::make_unique<T>(name, env, env.state); }

std::unique_ptr<Scenario> createScenario(const std::string &n, Env &env) {  // scenarios/init.hpp:28-56
    if (n == "towerbuilding") return make<TowerBuildingScenario>(n, env);
    if (n == "obstacleseasy") return make<ObstaclesEasyScenario>(n, env);
    if (n == "obstaclesmedium") return make<ObstaclesMediumScenario>(n, env);
    if (n == "obstacleshard") return make<ObstaclesHardScenario>(n, env);
    if (n == "obstacleswalls") return make<ObstaclesOnlyWallsScenario>(n, env);
    if (n == "obstaclessteps") return make<ObstaclesOnlyStepsScenario>(n, env);
    if (n == "obstacleslava") return make<ObstaclesOnlyLavaScenario>(n, env);
    if (n == "collect") return make<CollectScenario>(n, env);
    if (n == "sokoban") return make<SokobanScenario>(n, env);
    if (n == "rearrange") return make<RearrangeScenario>(n, env);
    if (n == "hexexplore") return make<HexExploreScenario>(n, env);
    if (n == "hexmemory") return make<HexMemoryScenario>(n, env);
    return nullptr;
}

uint32_t bits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

void setPoses(Handle &h, const float *poses) {  // per agent: objectT[16] cameraLocal[16] onGround
    for (int i = 0; i < h.env.numAgents; ++i) {
        auto *a = static_cast<DefaultKinematicAgent *>(h.env.state.agents[size_t(i)]);
        const float *p = poses + i * 33;
        a->setTransformation(Magnum::Matrix4::from(p));
        a->cameraObject->setTransformation(Magnum::Matrix4::from(p + 16));
        a->grounded = p[32] != 0.0f;
    }
}

}  // namespace

extern "C" {

void *ref_scen_create(const char *name, int numAgents, const char **keys, const float *vals, int nparams) {
    setLogLevel(WARNING);  // the scenarios log every reset at INFO / DEBUG
    auto h = std::make_unique<Handle>(numAgents);
    h->scenario = createScenario(toLower(name), h->env);  // Env::Env, env.cpp:37-50
    if (!h->scenario) return nullptr;
    h->env.scenarioPtr = h->scenario.get();
    h->scenario->init();
    FloatParams fp;
    for (int i = 0; i < nparams; ++i) fp[keys[i]] = vals[i];
    h->scenario->setCustomParameters(fp);
    return h.release();
}
void ref_scen_destroy(void *p) { delete static_cast<Handle *>(p); }
void ref_scen_seed(void *p, int seedValue) { static_cast<Handle *>(p)->env.state.rng.seed((unsigned long)seedValue); }  // env.cpp:54-57

// Env::reset up to and including spawnAgents (env.cpp:58-72); out = per agent (startingPosition xyz, rotationRad) bit patterns
void ref_scen_reset_begin(void *p, unsigned mazeSeedXor, uint32_t *spawnOut) {
    auto &h = *static_cast<Handle *>(p);
    auto &st = h.env.state;
    standinBodies().clear();
    st.reset();
    auto seed = randRange(0, 1 << 30, st.rng);
    st.rng.seed((unsigned long)seed);
    g_mazeSeed = unsigned(seed) ^ mazeSeedXor;
    for (int t = int(DrawableType::First); t < int(DrawableType::NumTypes); ++t) h.drawables[DrawableType(t)].clear();
    h.scenario->reset();
    h.scenario->spawnAgents(st.agents);
    for (int i = 0; i < h.env.numAgents; ++i) {
        auto *a = static_cast<DefaultKinematicAgent *>(st.agents[size_t(i)]);
        spawnOut[4 * i + 0] = bits(a->startingPosition.x()), spawnOut[4 * i + 1] = bits(a->startingPosition.y());
        spawnOut[4 * i + 2] = bits(a->startingPosition.z()), spawnOut[4 * i + 3] = bits(a->rotationRad);
    }
}
// the rest of Env::reset (env.cpp:74-76) once the puppets are posed
void ref_scen_reset_end(void *p, const float *poses) {
    auto &h = *static_cast<Handle *>(p);
    setPoses(h, poses);
    h.scenario->addEpisodeDrawables(h.drawables);
    h.scenario->addEpisodeAgentsDrawables(h.drawables);
    h.scenario->addUIDrawables(h.drawables);
}

// Env::step (env.cpp:84-152) with the kinematics replaced by the given poses; teleportOut: per agent (flag, x, y, z)
void ref_scen_step(void *p, const int *actions, const float *poses, uint32_t *teleportOut) {
    auto &h = *static_cast<Handle *>(p);
    auto &st = h.env.state;
    std::fill(st.lastReward.begin(), st.lastReward.end(), 0.0f);
    for (int i = 0; i < h.env.numAgents; ++i) {
        st.currAction[size_t(i)] = Action(actions[i]);
        static_cast<DefaultKinematicAgent *>(st.agents[size_t(i)])->teleports.clear();
    }
    h.scenario->preStep();
    setPoses(h, poses);
    for (auto agent : st.agents) agent->updateTransform();
    h.scenario->step();
    st.currEpisodeSec += st.lastFrameDurationSec;
    h.scenario->updateUI();
    if (st.currEpisodeSec >= h.env.episodeLengthSec()) st.done = true;
    for (int i = 0; i < h.env.numAgents; ++i) st.currAction[size_t(i)] = Action::Idle;
    for (int i = 0; i < int(st.agents.size()); ++i) st.totalReward[size_t(i)] += st.lastReward[size_t(i)];
    ++st.numFrames;
    for (int i = 0; i < h.env.numAgents; ++i) {
        auto *a = static_cast<DefaultKinematicAgent *>(st.agents[size_t(i)]);
        teleportOut[4 * i] = uint32_t(a->teleports.size());
        const auto t = a->teleports.empty() ? Magnum::Vector3{0, 0, 0} : a->teleports.back();
        teleportOut[4 * i + 1] = bits(t.x()), teleportOut[4 * i + 2] = bits(t.y()), teleportOut[4 * i + 3] = bits(t.z());
    }
}

// Same layout as orc_scenario_dump (oracle/orc_api.cpp).  Floats as bit patterns.
int ref_scen_dump(void *p, uint32_t *out, int cap) {
    auto &h = *static_cast<Handle *>(p);
    auto &st = h.env.state;
    std::vector<uint32_t> o;
    o.push_back(bits(h.env.episodeLengthSec()));
    o.push_back(st.done ? 1u : 0u);
    o.push_back(bits(st.currEpisodeSec));
    o.push_back(uint32_t(h.env.numAgents));
    for (int i = 0; i < h.env.numAgents; ++i) {
        o.push_back(bits(st.lastReward[size_t(i)]));
        o.push_back(bits(st.totalReward[size_t(i)]));
        o.push_back(bits(h.scenario->trueObjective(i)));
    }
    size_t n = 0;
    for (auto &[type, list] : h.drawables) n += list.size();
    o.push_back(uint32_t(n));
    // the renderer's read: absoluteTransformationMatrix() of every drawable, mesh type major (v4r_env_renderer.cpp:267-279,319-335)
    for (auto &[type, list] : h.drawables)
        for (auto &d : list) {
            o.push_back(uint32_t(type));
            const auto c = d.color;
            o.push_back((uint32_t(c.r() * 255.0f + 0.5f) << 16) | (uint32_t(c.g() * 255.0f + 0.5f) << 8) | uint32_t(c.b() * 255.0f + 0.5f));
            const auto m = d.objectPtr->absoluteTransformationMatrix();
            for (int k = 0; k < 16; ++k) o.push_back(bits(m.data()[k]));
        }
    auto &bodies = standinBodies();
    o.push_back(uint32_t(bodies.size()));
    for (auto *b : bodies) {
        for (int k = 0; k < 3; ++k) o.push_back(bits(b->colliderOrigin[k]));
        for (int k = 0; k < 3; ++k) o.push_back(bits(b->colliderScaling[k]));
        o.push_back(bits(b->colliderAxisX.x())), o.push_back(bits(b->colliderAxisX.z()));
        o.push_back(b->colliding() ? 1u : 0u);
    }
    if (int(o.size()) > cap) return -int(o.size());
    std::copy(o.begin(), o.end(), out);
    return int(o.size());
}

}  // extern "C"
