// raytrace_2 [scene[.json]] [out.png] — headless drop-in for the reference binary (src/main.cpp:3-6 -> App::Run).
// The reference reads <SRC_PATH>/local/data/settings.json with SRC_PATH baked in at compile time
// (src/CMakeLists.txt:39, Paths.hpp:3-4); here the source root comes from $RAYTRACE2_ROOT (default: cwd).
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/rt2.h"

int main(int argc, char* argv[]) {
  const char* root_env = std::getenv("RAYTRACE2_ROOT");
  std::string root = root_env ? root_env : ".";
  std::string settings = root + "/local/data/settings.json";
  std::string data = root + "/data";
  int rc = rt2_app_run(argc, argv, settings.c_str(), data.c_str());
  if (rc != RT2_OK) {
    std::fprintf(stderr, "raytrace_2: %s\n", rt2_last_error());
    return 1;
  }
  return 0;
}
