// ORACLE tooling: forwards to the Bullet stand-in (mini_bullet.hpp), see there.
#pragma once
#include <mini_bullet.hpp>
