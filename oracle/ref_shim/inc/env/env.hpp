// ORACLE tooling.  Stand-in for the reference's env/env.hpp when compiling scenarios/platforms.hpp and
// scenarios/component_voxel_grid.hpp alone: the real header pulls in Bullet (physics.hpp), which this image does not have, while
// those two headers only need the FloatParams alias (src/libs/env/include/env/env.hpp:85), the NAMES Env / Env::EnvState (reset()
// signature of ScenarioComponent) and the reference's own voxel_state.hpp.  Everything else comes from the reference's headers.
#pragma once
#include <map>
#include <string>

#include <env/const.hpp>
#include <env/voxel_state.hpp>

namespace Megaverse {
using FloatParams = std::map<std::string, float>;
class Env { public: struct EnvState {}; };
}
