// ORACLE tooling -- NOT product code.  Companion of the reference's pybind module bindings/megaverse.cpp when it is compiled in place
// (oracle/_ref/pyref/megaverse*.so, `make -C oracle pyref`) on the Bullet stand-in with null renderers: MegaverseGym::seed's master
// stream, setActions' encoding of the six Discrete heads, VectorEnv, getLastRewards' ordering, isDone / trueObjective and the reward
// shaping accessors are then the reference's own code, reachable from Python exactly as megaverse/megaverse_env.py reaches them.
// Holds the two substitutions of env_shim.cpp (seedable maze generator, correctly rounded sinf / cosf) for this module.
#include <cmath>
#include <deque>

#include <mazes/spanningtreealgorithm.h>

#include <null_renderer.hpp>

extern "C" {
float sinf(float x) noexcept { return float(std::sin(double(x))); }
float cosf(float x) noexcept { return float(std::cos(double(x))); }
void sincosf(float x, float *s, float *c) noexcept { *s = float(std::sin(double(x))); *c = float(std::cos(double(x))); }
}

static std::deque<unsigned> g_mazeSeedQueue;
SpanningtreeAlgorithm::SpanningtreeAlgorithm() {
    unsigned seed = 0;
    if (!g_mazeSeedQueue.empty()) { seed = g_mazeSeedQueue.front(); g_mazeSeedQueue.pop_front(); }
    generator = std::mt19937(seed);
}

namespace Megaverse {
static unsigned predicted(Env &env) {
    Rng copy = env.getRng();
    return unsigned(randRange(0, 1 << 30, copy)) ^ 0x6D617A65u;  // the oracle's derivation (orc_env.hpp hexMazeReset)
}
void standinQueueMazeSeed(Env &env) { g_mazeSeedQueue.push_back(predicted(env)); }
void standinQueueMazeSeedsForReset(Envs &envs) {
    g_mazeSeedQueue.clear();
    for (auto &e : envs) g_mazeSeedQueue.push_back(predicted(*e));
}
}  // namespace Megaverse
