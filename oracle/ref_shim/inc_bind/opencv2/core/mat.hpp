// ORACLE tooling: empty stand-in (bindings/megaverse.cpp includes OpenCV's Mat header without using it on the compared path).
#pragma once
