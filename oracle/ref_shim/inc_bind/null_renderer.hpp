// ORACLE tooling.  The renderers bindings/megaverse.cpp constructs (MagnumEnvRenderer, V4REnvRenderer) need OpenGL / Vulkan; these
// stand-ins draw nothing and hand out a zeroed frame.  They also carry the maze-seed substitution (see binding_shim.cpp): preDraw()
// runs right after every env's step, before VectorEnv::step resets the finished ones, which is the last moment the seed a hexagonal
// maze will be built from can be derived from that env's generator.
#pragma once
#include <cstdint>
#include <vector>

#include <env/env_renderer.hpp>
#include <env/vector_env.hpp>  // the real renderer headers bring this in for bindings/megaverse.cpp

namespace Megaverse {

void standinQueueMazeSeed(Env &env);        // binding_shim.cpp
void standinQueueMazeSeedsForReset(Envs &envs);

class NullEnvRenderer : public EnvRenderer {
public:
    NullEnvRenderer(Envs &envs, int w, int h, bool primary) : frame(size_t(w) * size_t(h) * 4, 0) {
        if (primary) standinQueueMazeSeedsForReset(envs);  // MegaverseGym::reset creates the renderer right before VectorEnv::reset
    }
    void reset(Env &, int) override {}
    void preDraw(Env &env, int) override { if (env.isDone()) standinQueueMazeSeed(env); }
    void draw(Envs &) override {}
    const uint8_t *getObservation(int, int) const override { return frame.data(); }
    Overview *getOverview() override { return nullptr; }
private:
    std::vector<uint8_t> frame;
};

}  // namespace Megaverse
