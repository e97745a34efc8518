// oracle shim (test infrastructure): minimal stand-in for SDL2's event header; only the members that
// src/cpu_raytrace/RayTracer.cpp:72-78 touches.  The live window is out of scope.
#pragma once
enum { SDL_WINDOWEVENT_RESIZED = 5 };
struct SDL_Event {
  struct { int event, data1, data2; } window;
};
