// ORACLE tooling: what magnum-integration's cmake would generate from configure.h.cmake for a static build.
#pragma once
#define MAGNUM_BULLETINTEGRATION_BUILD_STATIC
