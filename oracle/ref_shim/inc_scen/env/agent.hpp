// ORACLE tooling.  Stand-in for the reference's env/agent.hpp: the agent classes' stand-ins live in env.hpp.
#pragma once
#include <env/env.hpp>
