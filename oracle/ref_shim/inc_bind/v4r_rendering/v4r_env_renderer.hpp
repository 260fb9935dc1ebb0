// ORACLE tooling: stand-in for the Vulkan renderer's header (constructor signature of v4r_env_renderer.hpp), see null_renderer.hpp.
#pragma once
#include <null_renderer.hpp>
namespace Megaverse {
class V4REnvRenderer : public NullEnvRenderer {
public:
    V4REnvRenderer(Envs &envs, int w, int h, V4REnvRenderer *previous, bool) : NullEnvRenderer(envs, w, h, previous == nullptr) {}
};
}
