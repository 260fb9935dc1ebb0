// ORACLE tooling.  Empty stand-in for the reference's env/physics.hpp (Bullet wrappers); RigidBody's stand-in lives in env.hpp.
#pragma once
