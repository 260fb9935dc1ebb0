// ORACLE tooling: stand-in for Magnum's GL default framebuffer (no GL context here).  agent.cpp:37 only reads its viewport size to
// size the camera feature's viewport, which nothing on the compared path uses.
#pragma once
#include <Magnum/Magnum.h>
#include <Magnum/Math/Range.h>
namespace Magnum { namespace GL {
struct DefaultFramebufferStandIn { Range2Di viewport() const { return Range2Di{{0, 0}, {128, 72}}; } };
static DefaultFramebufferStandIn defaultFramebuffer;
}}
