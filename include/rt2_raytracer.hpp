// include/rt2_raytracer.hpp — header-only C++ adapter over the C ABI (include/rt2.h) with the method names of the reference's
// renderer, so an App-like driver swaps `raytrace2::cpu::RayTracer` for `raytrace2::b200::RayTracer` and keeps its call sites.
//
//   reference                                         (file:line, reference root)          adapter
//   cpu::RayTracer::Update(const Scene&)              src/cpu_raytrace/RayTracer.cpp:55     Update(scene)
//   cpu::RayTracer::OnResize(glm::ivec2)              RayTracer.cpp:87                      OnResize(w, h)
//   cpu::RayTracer::Reset()                           RayTracer.cpp:49                      Reset()
//   cpu::RayTracer::NonConvertedPixels()              RayTracer.cpp:105                     NonConvertedPixels()
//   cpu::RayTracer::Pixels()                          RayTracer.hpp:22                      Pixels()
//   cpu::RayTracer::FrameIdx() / Dims()               RayTracer.hpp:23,29                   FrameIdx() / Dims()
//   serialize::SceneLoader::LoadScene(path)           src/Serialize.hpp:21-22               SceneLoader::LoadScene(path)
//   util::WriteImage(pixels, w, h, path, png)         src/Util.hpp:11-12                    WriteImage(...)
//
// GLM-free on purpose (the reference's vec3 / ivec2 are replaced by std::array); errors become exceptions of type
// std::runtime_error carrying rt2_last_error().
#pragma once
#include <array>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "rt2.h"

namespace raytrace2::b200 {

inline void Check(int rc) {
  if (rc != RT2_OK) throw std::runtime_error(std::string("rt2: ") + rt2_last_error());
}

// ≡ cpu::Scene as App::Run holds it after loading (src/App.cpp:117-126): flattened + BVH built.
class Scene {
 public:
  Scene() = default;
  explicit Scene(rt2_scene* h) : h_(h) {}
  Scene(Scene&& o) noexcept : h_(std::exchange(o.h_, nullptr)) {}
  Scene& operator=(Scene&& o) noexcept {
    if (this != &o) {
      rt2_scene_destroy(h_);
      h_ = std::exchange(o.h_, nullptr);
    }
    return *this;
  }
  Scene(const Scene&) = delete;
  Scene& operator=(const Scene&) = delete;
  ~Scene() { rt2_scene_destroy(h_); }
  [[nodiscard]] rt2_scene* Handle() const { return h_; }
  [[nodiscard]] std::array<int, 2> dims() const {
    rt2_scene_desc d;
    Check(rt2_scene_get_desc(h_, &d));
    return {d.width, d.height};
  }

 private:
  rt2_scene* h_{nullptr};
};

struct SceneLoader {
  std::string data_dir;  // GET_PATH("data/") of the reference (Paths.hpp:3)
  uint64_t perlin_seed{0};
  // nullopt where the reference's loader returns nullopt / would throw
  [[nodiscard]] std::optional<Scene> LoadScene(const std::string& filepath) const {
    rt2_scene* h = nullptr;
    if (rt2_scene_load(filepath.c_str(), data_dir.empty() ? nullptr : data_dir.c_str(), perlin_seed, &h) != RT2_OK) return std::nullopt;
    return Scene(h);
  }
};

class RayTracer {
 public:
  size_t max_depth{50};     // RayTracer::max_depth (RayTracer.hpp:32); applied at Init
  int num_samples{1};       // Camera::SetSamplesPerPixel (App.cpp:129): stratification grid
  int device{0};
  int n_gpus{-1};           // -1: every GPU of the box behind this one object (rt2_config.n_gpus); 1: only `device`
  uint64_t seed{0x5EED};

  RayTracer() = default;
  RayTracer(const RayTracer&) = delete;
  RayTracer& operator=(const RayTracer&) = delete;
  ~RayTracer() { rt2_destroy(r_); }

  // Binds the renderer to a scene (the reference passes `const Scene&` to every Update and a Camera* once, App.cpp:130).
  void Init(const Scene& scene, int frame_offset = 0, int frame_stride = 1) {
    rt2_destroy(r_);
    r_ = nullptr;
    rt2_config cfg{};
    cfg.device = device;
    cfg.n_gpus = n_gpus;
    cfg.samples_per_pixel = num_samples;
    cfg.max_depth = static_cast<int32_t>(max_depth);
    cfg.frame_offset = frame_offset;
    cfg.frame_stride = frame_stride;
    cfg.seed = seed;
    Check(rt2_create(scene.Handle(), &cfg, &r_));
  }
  // +1 sample per pixel (RayTracer.cpp:55-70).  Cheap: frames are collected into wavefront batches and traced when a batch is
  // full or when pixels are read, so the reference's one-Update-per-sample loop (App.cpp:244-246) needs no change.
  void Update(const Scene&) { Check(rt2_update(r_, 1)); }
  void Update(const Scene&, uint32_t n_frames) { Check(rt2_update(r_, n_frames)); }  // n x Update in one call
  void OnResize(int w, int h) { Check(rt2_resize(r_, w, h)); }
  void Reset() { Check(rt2_reset(r_)); }
  [[nodiscard]] size_t FrameIdx() const {
    uint64_t f = 0;
    Check(rt2_frame_idx(r_, &f));
    return static_cast<size_t>(f);
  }
  [[nodiscard]] std::array<int, 2> Dims() const {
    int32_t w = 0, h = 0;
    Check(rt2_dims(r_, &w, &h));
    return {w, h};
  }
  // float RGB mean, row 0 = bottom of the image (RayTracer.cpp:105-112)
  [[nodiscard]] std::vector<std::array<float, 3>> NonConvertedPixels() const {
    auto d = Dims();
    std::vector<std::array<float, 3>> px(static_cast<size_t>(d[0]) * d[1]);
    Check(rt2_read_mean_rgb32f(r_, reinterpret_cast<float*>(px.data())));
    return px;
  }
  // RGBA8 preview, linear (RayTracer.cpp:16-18,65-66)
  [[nodiscard]] std::vector<std::array<uint8_t, 4>> Pixels() const {
    auto d = Dims();
    std::vector<std::array<uint8_t, 4>> px(static_cast<size_t>(d[0]) * d[1]);
    Check(rt2_read_rgba8(r_, reinterpret_cast<uint8_t*>(px.data())));
    return px;
  }
  [[nodiscard]] rt2_stats Stats() const {
    rt2_stats s{};
    Check(rt2_get_stats(r_, &s));
    return s;
  }

 private:
  rt2_renderer* r_{nullptr};
};

inline void WriteImage(const std::vector<std::array<float, 3>>& pixels, int width, int height, const std::string& out_path, bool png = true) {
  Check(rt2_write_image(reinterpret_cast<const float*>(pixels.data()), width, height, out_path.c_str(), png ? 1 : 0));
}

}  // namespace raytrace2::b200
