// examples/progressive_view.cpp — the LIVE branch of the reference's App::Run (src/App.cpp:176-242) without a window: the
// render loop, its frame counter, the RGBA8 preview upload and the two UI controls, written against include/rt2_raytracer.hpp.
//
//   reference (src/App.cpp)                                              here
//   :196-199  if (!done || !render_once) cpu_tracer_.Update(scene)       tracer.Update(scene) once per loop iteration
//   :217      ImGui::Text("Frame Count %i", FrameIdx())                  printed with every preview
//   :220-227  "Load Scene" button -> LoadScene + cpu_tracer_.Reset()     script event  load:<scene.json>
//   :229      cpu_tracer_.OnImGui(): "Reset" button -> Reset()           script event  reset
//   RayTracer.cpp:72-77 window resize -> OnResize(dims)                  script event  resize:<w>x<h>
//   :234-236  glTextureSubImage2D(..., cpu_tracer_.Pixels().data())      Pixels() every `preview_every` iterations -> PPM file
//                                                                        (GL_FRAMEBUFFER_SRGB at :237 = the viewer's job)
//
// Update() only collects frames; they are traced in wavefront batches when Pixels() is read, so a preview every K iterations
// costs one batch of K frames + one 1.4 MB read-back.  Events are given as "<iteration>:<event>" arguments.
//
//   g++ -std=c++17 -Iinclude examples/progressive_view.cpp -Lraytrace2_b200/lib -lraytrace2_b200 -Wl,-rpath,$PWD/raytrace2_b200/lib -o /tmp/progressive_view
//   /tmp/progressive_view data/cornell_original_test.json /tmp/preview 64 16 24:reset 40:resize:300x200 48:load:data/cornell_box4.json
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <string>
#include <vector>

#include "rt2_raytracer.hpp"

namespace b200 = raytrace2::b200;

// RGBA8, row 0 = bottom (RayTracer.cpp:97-102) -> binary PPM, top row first
static void WritePreview(const std::vector<std::array<uint8_t, 4>>& px, int w, int h, const std::string& path) {
  std::ofstream f(path, std::ios::binary);
  f << "P6\n" << w << " " << h << "\n255\n";
  for (int y = h - 1; y >= 0; y--)
    for (int x = 0; x < w; x++) f.write(reinterpret_cast<const char*>(px[static_cast<size_t>(y) * w + x].data()), 3);
}

int main(int argc, char** argv) {
  if (argc < 5) {
    std::fprintf(stderr, "usage: %s scene.json out_prefix iterations preview_every [<iteration>:reset|resize:<w>x<h>|load:<scene.json>]...\n", argv[0]);
    return 2;
  }
  const std::string prefix = argv[2];
  const int iterations = std::atoi(argv[3]), preview_every = std::atoi(argv[4]) > 0 ? std::atoi(argv[4]) : 1;
  std::multimap<int, std::string> events;
  for (int i = 5; i < argc; i++) {
    const char* colon = std::strchr(argv[i], ':');
    if (!colon) {
      std::fprintf(stderr, "bad event '%s'\n", argv[i]);
      return 2;
    }
    events.emplace(std::atoi(argv[i]), std::string(colon + 1));
  }
  b200::SceneLoader loader;
  auto scene_opt = loader.LoadScene(argv[1]);
  if (!scene_opt.has_value()) {
    std::fprintf(stderr, "Failed to parse Scene: %s. %s\n", rt2_last_error(), argv[1]);
    return 1;
  }
  b200::Scene scene = std::move(scene_opt.value());
  b200::RayTracer tracer;
  tracer.max_depth = 50;
  tracer.num_samples = iterations;
  tracer.Init(scene);
  int previews = 0;
  for (int it = 0; it < iterations; it++) {
    auto range = events.equal_range(it);
    for (auto e = range.first; e != range.second; ++e) {
      const std::string& ev = e->second;
      if (ev == "reset") {
        tracer.Reset();  // RayTracer::OnImGui "Reset" (RayTracer.cpp:79-85)
        std::printf("[%d] Reset -> Frame Count %zu\n", it, tracer.FrameIdx());
      } else if (ev.rfind("resize:", 0) == 0) {
        int w = 0, h = 0;
        if (std::sscanf(ev.c_str() + 7, "%dx%d", &w, &h) != 2) return 2;
        tracer.OnResize(w, h);  // SDL_WINDOWEVENT_RESIZED (RayTracer.cpp:72-77): camera dims, realloc, frame_idx_ = 0
        std::printf("[%d] OnResize %dx%d -> Frame Count %zu\n", it, w, h, tracer.FrameIdx());
      } else if (ev.rfind("load:", 0) == 0) {
        // "Load Scene" (App.cpp:220-227): a failed load keeps the old scene; a good one replaces it and resets the tracer
        auto next = loader.LoadScene(ev.substr(5));
        if (next.has_value()) {
          scene = std::move(next.value());
          tracer.Init(scene);
          std::printf("[%d] Load Scene %s -> Frame Count %zu\n", it, ev.c_str() + 5, tracer.FrameIdx());
        } else {
          std::printf("[%d] Load Scene %s failed: %s (scene kept)\n", it, ev.c_str() + 5, rt2_last_error());
        }
      }
    }
    tracer.Update(scene);  // App.cpp:196-199
    if ((it + 1) % preview_every == 0 || it + 1 == iterations) {
      auto d = tracer.Dims();
      const std::string path = prefix + "_" + std::to_string(previews++) + ".ppm";
      WritePreview(tracer.Pixels(), d[0], d[1], path);  // App.cpp:234-236
      std::printf("[%d] Frame Count %zu  preview %s (%dx%d)\n", it, tracer.FrameIdx(), path.c_str(), d[0], d[1]);
    }
  }
  return 0;
}
