// Tuning knobs of A/B experiments.  The shipped library is built WITHOUT RT2_EXPERIMENTS: every knob is then the constant
// that was measured best (profiles/), and no environment variable influences the renderer.  `make EXPERIMENTS=1` builds
// libraytrace2_b200_exp.so, in which RT2_* environment variables override them (tools/exp_*.sh, tools/bvh_sweep.sh).
#pragma once
#include <cstdlib>

namespace rt2 {

inline long TuneInt(const char* name, long def) {
#ifdef RT2_EXPERIMENTS
  if (const char* e = std::getenv(name)) return std::atol(e);
#else
  (void)name;
#endif
  return def;
}
inline double TuneFloat(const char* name, double def) {
#ifdef RT2_EXPERIMENTS
  if (const char* e = std::getenv(name)) return std::atof(e);
#else
  (void)name;
#endif
  return def;
}

}  // namespace rt2
