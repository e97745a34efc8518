// C ABI of libraytrace2_b200.so (include/rt2.h): thin, exception-free wrappers over the host scene compiler
// (host/scene_host.cpp), the wavefront renderer (device/rt_kernels.cu) and the image writer (host/image_out.cpp).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <new>
#include <string>
#include <vector>

#include <sys/stat.h>

#include "../../include/rt2.h"
#include "device/rt_multi.hpp"
#include "host/image_out.hpp"
#include "host/scene_host.hpp"

struct rt2_scene {
  rt2::HostScene host;
};
struct rt2_renderer {
  rt2::MultiRenderer impl;  // one wavefront renderer per GPU behind the handle (device/rt_multi.hpp)
};

namespace {
thread_local std::string g_last_error;
int Fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
}  // namespace

extern "C" {

const char* rt2_last_error(void) { return g_last_error.c_str(); }
int rt2_abi_version(void) { return RT2_ABI_VERSION; }
int rt2_device_count(void) { return rt2::DeviceCount(); }
int rt2_measure_fp32_peak(int32_t device, double* tflops) {
  if (!tflops) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  std::string err;
  int rc = rt2::MeasureFp32Peak(device, tflops, &err);
  return rc == RT2_OK ? RT2_OK : Fail(rc, err);
}
int rt2_measure_l2_bandwidth(int32_t device, double* gbs) {
  if (!gbs) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  std::string err;
  int rc = rt2::MeasureL2Bandwidth(device, gbs, &err);
  return rc == RT2_OK ? RT2_OK : Fail(rc, err);
}

// ---- scene -----------------------------------------------------------------------------------------------------
int rt2_scene_load(const char* json_path, const char* data_dir, uint64_t perlin_seed, rt2_scene** out) {
  if (!json_path || !out) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  try {
    auto* s = new rt2_scene;
    std::string err;
    int rc = rt2::LoadSceneFile(json_path, data_dir ? data_dir : "", perlin_seed, &s->host, &err);
    if (rc != RT2_OK) {
      delete s;
      return Fail(rc, err);
    }
    *out = s;
    return RT2_OK;
  } catch (const std::exception& e) {
    return Fail(RT2_ERR_PARSE, std::string("scene load failed: ") + e.what());
  }
}

int rt2_scene_load_string(const char* json_text, const char* data_dir, uint64_t perlin_seed, rt2_scene** out) {
  if (!json_text || !out) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  try {
    auto* s = new rt2_scene;
    std::string err;
    int rc = rt2::LoadSceneString(json_text, data_dir ? data_dir : ".", perlin_seed, &s->host, &err);
    if (rc != RT2_OK) {
      delete s;
      return Fail(rc, err);
    }
    *out = s;
    return RT2_OK;
  } catch (const std::exception& e) {
    return Fail(RT2_ERR_PARSE, std::string("scene load failed: ") + e.what());
  }
}

int rt2_scene_synthetic_spheres(uint32_t n_spheres, uint64_t seed, int32_t width, int32_t height, int32_t build_host_bvh,
                                rt2_scene** out) {
  if (!out) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  try {
    auto* s = new rt2_scene;
    std::string err;
    int rc = rt2::MakeSyntheticSpheres(n_spheres, seed, width, height, build_host_bvh != 0, &s->host, &err);
    if (rc != RT2_OK) {
      delete s;
      return Fail(rc, err);
    }
    *out = s;
    return RT2_OK;
  } catch (const std::exception& e) {
    return Fail(RT2_ERR_INVALID_ARG, std::string("synthetic scene failed: ") + e.what());
  }
}

void rt2_scene_destroy(rt2_scene* scene) { delete scene; }

int rt2_scene_get_desc(const rt2_scene* scene, rt2_scene_desc* out) {
  if (!scene || !out) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  scene->host.FillDesc(out);
  return RT2_OK;
}

int rt2_scene_set_dims(rt2_scene* scene, int32_t width, int32_t height) {
  if (!scene || width <= 0 || height <= 0) return Fail(RT2_ERR_INVALID_ARG, "invalid dims");
  scene->host.width = width;
  scene->host.height = height;
  scene->host.UpdateCamera();
  return RT2_OK;
}

int rt2_scene_set_perlin(rt2_scene* scene, uint32_t perlin_idx, const int32_t* perm_x, const int32_t* perm_y,
                         const int32_t* perm_z, const float* vec) {
  if (!scene || !perm_x || !perm_y || !perm_z || !vec) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  if (perlin_idx >= scene->host.perlin.size()) return Fail(RT2_ERR_INVALID_ARG, "perlin index out of range");
  rt2_perlin& p = scene->host.perlin[perlin_idx];
  for (int i = 0; i < 256; i++) {
    p.perm_x[i] = perm_x[i] & 255;
    p.perm_y[i] = perm_y[i] & 255;
    p.perm_z[i] = perm_z[i] & 255;
    for (int k = 0; k < 3; k++) p.vec[i][k] = vec[i * 3 + k];
    p.vec[i][3] = 0;
  }
  return RT2_OK;
}

int rt2_scene_get_perlin(const rt2_scene* scene, uint32_t perlin_idx, int32_t* perm_x, int32_t* perm_y, int32_t* perm_z,
                         float* vec) {
  if (!scene || !perm_x || !perm_y || !perm_z || !vec) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  if (perlin_idx >= scene->host.perlin.size()) return Fail(RT2_ERR_INVALID_ARG, "perlin index out of range");
  const rt2_perlin& p = scene->host.perlin[perlin_idx];
  for (int i = 0; i < 256; i++) {
    perm_x[i] = p.perm_x[i];
    perm_y[i] = p.perm_y[i];
    perm_z[i] = p.perm_z[i];
    for (int k = 0; k < 3; k++) vec[i * 3 + k] = p.vec[i][k];
  }
  return RT2_OK;
}

int rt2_scene_span1_flags(const rt2_scene* scene, uint8_t* flags) {
  if (!scene || !flags) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  std::memcpy(flags, scene->host.span1_flags.data(), scene->host.span1_flags.size());
  return RT2_OK;
}

// ---- renderer --------------------------------------------------------------------------------------------------
int rt2_create(const rt2_scene* scene, const rt2_config* cfg, rt2_renderer** out) {
  if (!scene || !cfg || !out) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  rt2_renderer* r = new (std::nothrow) rt2_renderer;
  if (!r) return Fail(RT2_ERR_INVALID_ARG, "out of host memory");
  int rc = r->impl.Init(scene->host, *cfg);
  if (rc != RT2_OK) {
    std::string msg = r->impl.Error();
    delete r;
    return Fail(rc, msg);
  }
  *out = r;
  return RT2_OK;
}

void rt2_destroy(rt2_renderer* r) { delete r; }

#define RT2_FORWARD(expr)                                   \
  if (!r) return Fail(RT2_ERR_INVALID_ARG, "null renderer"); \
  {                                                         \
    int rc__ = (expr);                                      \
    if (rc__ != RT2_OK) return Fail(rc__, r->impl.Error()); \
    return RT2_OK;                                          \
  }

int rt2_upload_scene(rt2_renderer* r, const rt2_scene* scene) {
  if (!scene) return Fail(RT2_ERR_INVALID_ARG, "null scene");
  RT2_FORWARD(r->impl.UploadScene(scene->host))
}
int rt2_resize(rt2_renderer* r, int32_t width, int32_t height) { RT2_FORWARD(r->impl.Resize(width, height)) }
int rt2_reset(rt2_renderer* r) { RT2_FORWARD(r->impl.Reset()) }
int rt2_update(rt2_renderer* r, uint32_t n_frames) { RT2_FORWARD(r->impl.Update(n_frames)) }
int rt2_flush(rt2_renderer* r) { RT2_FORWARD(r->impl.Flush()) }
int rt2_synchronize(rt2_renderer* r) { RT2_FORWARD(r->impl.Synchronize()) }
int rt2_frame_idx(const rt2_renderer* r, uint64_t* out) {
  if (!r || !out) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  *out = r->impl.FrameIdx();
  return RT2_OK;
}
int rt2_dims(const rt2_renderer* r, int32_t* width, int32_t* height) {
  if (!r || !width || !height) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  *width = r->impl.Width();
  *height = r->impl.Height();
  return RT2_OK;
}
int rt2_read_mean_rgb32f(rt2_renderer* r, float* dst) {
  if (!dst) return Fail(RT2_ERR_INVALID_ARG, "null destination");
  RT2_FORWARD(r->impl.ReadMean(dst))
}
int rt2_read_rgba8(rt2_renderer* r, uint8_t* dst) {
  if (!dst) return Fail(RT2_ERR_INVALID_ARG, "null destination");
  RT2_FORWARD(r->impl.ReadRGBA8(dst))
}
int rt2_read_accum(rt2_renderer* r, float* sum, float* sumsq) { RT2_FORWARD(r->impl.ReadAccum(sum, sumsq)) }
int rt2_write_accum(rt2_renderer* r, const float* sum, const float* sumsq, uint64_t frames) {
  RT2_FORWARD(r->impl.WriteAccum(sum, sumsq, frames))
}
// The next four calls are the plumbing of the one-process-per-GPU harness (external NCCL reduce, CUDA IPC read-out); a
// multi-GPU handle reduces over peer memory by itself and rejects them.
#define RT2_SINGLE(what)                                               \
  if (!r) return Fail(RT2_ERR_INVALID_ARG, "null renderer");           \
  rt2::Renderer* one = r->impl.Single(what);                           \
  if (!one) return Fail(RT2_ERR_UNSUPPORTED, r->impl.Error());
#define RT2_SINGLE_RC(expr)                             \
  {                                                     \
    int rc__ = (expr);                                  \
    if (rc__ != RT2_OK) return Fail(rc__, one->Error()); \
    return RT2_OK;                                      \
  }
int rt2_accum_ipc_handle(rt2_renderer* r, uint8_t* handle) {
  if (!handle) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  RT2_SINGLE("rt2_accum_ipc_handle")
  RT2_SINGLE_RC(one->AccumIpcHandle(handle))
}
int rt2_resolve_peers(rt2_renderer* r, const uint8_t* handles, uint32_t n_ranks, uint32_t self_rank, uint64_t total_frames, float* dst_mean_rgb,
                      uint8_t* dst_rgba8) {
  RT2_SINGLE("rt2_resolve_peers")
  RT2_SINGLE_RC(one->ResolvePeers(handles, n_ranks, self_rank, total_frames, dst_mean_rgb, dst_rgba8))
}
int rt2_accum_device_ptr(rt2_renderer* r, void** ptr, size_t* n_floats) {
  if (!ptr || !n_floats) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  RT2_SINGLE("rt2_accum_device_ptr")
  RT2_SINGLE_RC(one->AccumDevicePtr(ptr, n_floats))
}
int rt2_set_frame_idx(rt2_renderer* r, uint64_t frames) { RT2_FORWARD(r->impl.SetFrameIdx(frames)) }
// Fixed-ray / fixed-point hooks and BVH inspection run on the first GPU (every replica holds the same scene).
#define RT2_FIRST_RC(expr)                                           \
  if (!r) return Fail(RT2_ERR_INVALID_ARG, "null renderer");         \
  {                                                                  \
    int rc__ = (expr);                                               \
    if (rc__ != RT2_OK) return Fail(rc__, r->impl.First().Error());  \
    return RT2_OK;                                                   \
  }
int rt2_intersect(rt2_renderer* r, const float* rays, size_t n, float tmin, float tmax, int skip_media, rt2_hit* out) {
  if (n > 0 && (!rays || !out)) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  RT2_FIRST_RC(r->impl.First().Intersect(rays, n, tmin, tmax, skip_media, out))
}
int rt2_texture_value(rt2_renderer* r, uint32_t tex_idx, const float* points, const float* uv, size_t n, float* rgb) {
  if (n > 0 && (!points || !rgb)) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  RT2_FIRST_RC(r->impl.First().TextureValue(tex_idx, points, uv, n, rgb))
}
int rt2_read_bvh(rt2_renderer* r, rt2_bvh_node* nodes, size_t max_nodes, uint32_t* prim_refs, size_t max_refs, uint32_t* n_pairs,
                 uint32_t* n_refs, uint32_t* tlas_root) {
  if (!n_pairs || !n_refs || !tlas_root) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  RT2_FIRST_RC(r->impl.First().ReadBvh(nodes, max_nodes, prim_refs, max_refs, n_pairs, n_refs, tlas_root))
}
int rt2_get_stats(rt2_renderer* r, rt2_stats* out) {
  if (!out) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  RT2_FORWARD(r->impl.GetStats(out))
}
int rt2_read_queue_sizes(rt2_renderer* r, uint32_t* out, uint32_t max_bounces, uint32_t* n_bounces) {
  if (!out || !n_bounces) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  RT2_FIRST_RC(r->impl.First().QueueSizes(out, max_bounces, n_bounces))
}
int rt2_debug_counters(rt2_renderer* r, uint64_t* counters, int* enabled) {
  if (!counters || !enabled) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  RT2_FIRST_RC(r->impl.First().DebugCounters(counters, enabled))
}
int rt2_set_profiling(rt2_renderer* r, int enabled) {
  if (!r) return Fail(RT2_ERR_INVALID_ARG, "null renderer");
  r->impl.SetProfiling(enabled != 0);
  return RT2_OK;
}
int rt2_stream(rt2_renderer* r, void** stream) {
  if (!r || !stream) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  *stream = r->impl.First().Stream();
  return RT2_OK;
}

// ---- output ----------------------------------------------------------------------------------------------------
int rt2_write_image(const float* mean_rgb, int32_t width, int32_t height, const char* out_path, int png) {
  if (!mean_rgb || !out_path) return Fail(RT2_ERR_INVALID_ARG, "null argument");
  std::string err;
  if (!rt2::WriteImage(mean_rgb, width, height, out_path, png != 0, &err)) return Fail(RT2_ERR_IO, err);
  return RT2_OK;
}
int rt2_tonemap_rgb8(const float* mean_rgb, int32_t width, int32_t height, uint8_t* dst) {
  if (!mean_rgb || !dst || width <= 0 || height <= 0) return Fail(RT2_ERR_INVALID_ARG, "invalid argument");
  rt2::TonemapRGB8(mean_rgb, width, height, dst);
  return RT2_OK;
}

// ---- app: headless branch of App::Run (App.cpp:81-130,157,163-174,243-248) ------------------------------------------
int rt2_app_run(int argc_in, const char* const* argv_in, const char* settings_path, const char* data_dir) {
  rt2::AppSettings settings;
  std::string err;
  if (!settings_path) return Fail(RT2_ERR_INVALID_ARG, "settings path required");
  int rc = rt2::LoadAppSettings(settings_path, &settings, &err);
  if (rc != RT2_OK) return Fail(rc, err);
  std::string dd = data_dir ? data_dir : "data";
  // Extensions (not in the reference): --gpus N, --spp N, --max-depth N, --seed N; everything else is the reference's argv.
  std::vector<const char*> argv;
  int n_gpus = -1;  // every GPU of the box
  uint64_t seed = static_cast<uint64_t>(std::chrono::steady_clock::now().time_since_epoch().count());
  for (int i = 0; i < argc_in; i++) {
    const std::string a = argv_in[i] ? argv_in[i] : "";
    auto value = [&](long long* out) {
      if (i + 1 >= argc_in) return false;
      char* end = nullptr;
      *out = std::strtoll(argv_in[i + 1], &end, 10);
      if (end == argv_in[i + 1] || *end != 0) return false;
      i++;
      return true;
    };
    long long v = 0;
    if (i > 0 && (a == "--gpus" || a == "--spp" || a == "--max-depth" || a == "--seed")) {
      if (!value(&v) || (a != "--seed" && v < 1)) return Fail(RT2_ERR_INVALID_ARG, a + " needs a positive integer");
      if (a == "--gpus") n_gpus = static_cast<int>(v);
      else if (a == "--spp") settings.num_samples = static_cast<size_t>(v);
      else if (a == "--max-depth") settings.max_depth = static_cast<size_t>(v);
      else seed = static_cast<uint64_t>(v);
      continue;
    }
    argv.push_back(argv_in[i]);
  }
  const int argc = static_cast<int>(argv.size());
  // App.cpp:84-107
  std::string full_scene_path, filename;
  if (argc <= 1) {
    full_scene_path = dd + "/scene2.json";
    filename = "scene2";
  } else {
    full_scene_path = argv[1];
    const std::string suffix = ".json";
    if (full_scene_path.length() >= suffix.length() &&
        full_scene_path.compare(full_scene_path.length() - suffix.length(), suffix.length(), suffix) == 0) {
      filename = full_scene_path.substr(0, full_scene_path.length() - suffix.length());
    } else {
      filename = full_scene_path;
      full_scene_path += suffix;
    }
  }
  bool user_defined_output_path = false;
  std::string image_output_path;
  if (argc == 3) {
    user_defined_output_path = true;
    image_output_path = argv[2];
  }
  // App.cpp:108-113
  std::printf("Render window: %d\n", settings.render_window ? 1 : 0);
  std::printf("Render once: %d\n", settings.render_once ? 1 : 0);
  std::printf("Num Samples: %zu\n", settings.num_samples);
  std::printf("Max Depth: %zu\n", settings.max_depth);
  std::printf("Save Output: %d\n", settings.save_after_render_once ? 1 : 0);
  std::printf("Scene Path: %s\n", full_scene_path.c_str());
  if (settings.render_window) {
    std::fprintf(stderr, "note: render_window=true ignored — this backend is headless (no display on a B200 server)\n");
  }
  rt2_scene* scene = nullptr;
  rc = rt2_scene_load(full_scene_path.c_str(), dd.c_str(), static_cast<uint64_t>(std::time(nullptr)), &scene);
  if (rc != RT2_OK) {
    std::fprintf(stderr, "Failed to parse Scene: %s. %s\n", rt2_last_error(), full_scene_path.c_str());
    return rc;  // App.cpp:118-120: exit(1)
  }
  for (const std::string& w : scene->host.warnings) std::fprintf(stderr, "Scene warning: %s. %s\n", w.c_str(), full_scene_path.c_str());
  const int visible = rt2_device_count();
  if (n_gpus > visible) {
    rt2_scene_destroy(scene);
    return Fail(RT2_ERR_INVALID_ARG, "--gpus " + std::to_string(n_gpus) + ": only " + std::to_string(visible) + " CUDA device(s) visible");
  }
  rt2_config cfg{};
  cfg.device = 0;
  cfg.n_gpus = n_gpus;
  cfg.samples_per_pixel = static_cast<int32_t>(settings.num_samples);
  cfg.max_depth = static_cast<int32_t>(settings.max_depth);
  cfg.seed = seed;
  rt2_renderer* r = nullptr;
  rc = rt2_create(scene, &cfg, &r);
  if (rc != RT2_OK) {
    std::fprintf(stderr, "rt2_create failed: %s\n", rt2_last_error());
    rt2_scene_destroy(scene);
    return rc;
  }
  // App.cpp:243-248: `if (frame_idx < num_samples) Update(scene)` once per sample, then write the image.  The calls only
  // collect frames; the GPUs trace them in wavefront batches (rt2_update).
  const auto t0 = std::chrono::steady_clock::now();
  for (size_t i = 0; rc == RT2_OK && i < settings.num_samples; i++) rc = rt2_update(r, 1);
  int w = 0, h = 0;
  rt2_dims(r, &w, &h);
  std::vector<float> mean(static_cast<size_t>(w) * h * 3);
  if (rc == RT2_OK) rc = rt2_read_mean_rgb32f(r, mean.data());
  const double wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (rc == RT2_OK) {
    rt2_stats st{};
    rt2_get_stats(r, &st);
    if (st.gpu_ms_total > 0) {
      std::printf("Rendered %llu paths, %llu rays on %u GPU(s) in %.3f s (render + reduce + read-back; %.1f Mrays/s)\n",
                  static_cast<unsigned long long>(st.paths), static_cast<unsigned long long>(st.rays), st.n_gpus, wall_s,
                  st.rays / wall_s * 1e-6);
    }
    // App.cpp:163-174
    if (!user_defined_output_path) {
      std::string out_dir = "local/output/";
      mkdir("local", 0755);
      mkdir(out_dir.c_str(), 0755);
      time_t now = time(nullptr);
      struct tm tstruct = *localtime(&now);
      char buf[80];
      strftime(buf, sizeof(buf), "%Y-%m-%d.%X", &tstruct);
      std::string base = filename;
      size_t slash = base.find_last_of('/');
      if (slash != std::string::npos) base = base.substr(slash + 1);
      image_output_path = out_dir + base + "_" + buf + ".png";
    }
    std::printf("Writing image: %s\n", image_output_path.c_str());
    rc = rt2_write_image(mean.data(), w, h, image_output_path.c_str(), 1);
  }
  if (rc != RT2_OK) std::fprintf(stderr, "render failed: %s\n", rt2_last_error());
  rt2_destroy(r);
  rt2_scene_destroy(scene);
  return rc;
}

}  // extern "C"
