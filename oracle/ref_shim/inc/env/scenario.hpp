// ORACLE tooling.  Stand-in for the reference's env/scenario.hpp: scenario_component.hpp only stores a Scenario reference.
#pragma once
#include <env/env.hpp>

namespace Megaverse { class Scenario {}; }
