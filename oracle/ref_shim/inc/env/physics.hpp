// ORACLE tooling.  Empty stand-in for the reference's env/physics.hpp (Bullet wrappers): voxel_state.hpp includes it but uses
// nothing from it.
#pragma once
