// oracle shim (test infrastructure): no-op ImGui calls used by src/cpu_raytrace/RayTracer.cpp:80-86.
#pragma once
namespace ImGui {
inline bool Begin(const char*) { return true; }
inline bool Button(const char*) { return false; }
inline void End() {}
}  // namespace ImGui
