// examples/headless_app.cpp — the headless branch of the reference's App::Run (src/App.cpp:115-130,157,243-248) written
// against the adapter in include/rt2_raytracer.hpp: what the reference's call sites look like after the swap.
//   g++ -std=c++17 -Iinclude examples/headless_app.cpp -Lraytrace2_b200/lib -lraytrace2_b200 -Wl,-rpath,$PWD/raytrace2_b200/lib -o /tmp/headless_app
//   /tmp/headless_app data/cornell_original_test.json out.png 64
#include <cstdio>
#include <cstdlib>

#include "rt2_raytracer.hpp"

int main(int argc, char** argv) {
  if (argc < 3) {
    std::fprintf(stderr, "usage: %s scene.json out.png [num_samples]\n", argv[0]);
    return 2;
  }
  const int num_samples = argc > 3 ? std::atoi(argv[3]) : 16;
  raytrace2::b200::SceneLoader loader;
  auto scene_opt = loader.LoadScene(argv[1]);  // App.cpp:116-120
  if (!scene_opt.has_value()) {
    std::fprintf(stderr, "Failed to parse Scene: %s. %s\n", rt2_last_error(), argv[1]);
    return 1;
  }
  raytrace2::b200::Scene& scene = scene_opt.value();
  raytrace2::b200::RayTracer tracer;
  tracer.max_depth = 50;             // App.cpp:128
  tracer.num_samples = num_samples;  // App.cpp:129
  tracer.Init(scene);                // App.cpp:130 + OnResize(initial_dims), App.cpp:157
  for (int i = 0; i < num_samples; i++) tracer.Update(scene);  // App.cpp:244-246
  auto dims = tracer.Dims();
  std::printf("Writing image: %s\n", argv[2]);  // App.cpp:171
  raytrace2::b200::WriteImage(tracer.NonConvertedPixels(), dims[0], dims[1], argv[2]);
  return 0;
}
