// ORACLE tooling: stand-in for the OpenGL renderer's header (constructor signature of magnum_env_renderer.hpp), see null_renderer.hpp.
#pragma once
#include <null_renderer.hpp>
namespace Megaverse {
class MagnumEnvRenderer : public NullEnvRenderer {
public:
    MagnumEnvRenderer(Envs &envs, int w, int h) : NullEnvRenderer(envs, w, h, false) {}
};
}
