// oracle/ref_driver.cpp — TEST INFRASTRUCTURE.  Never linked into, imported by, or shipped with the product.
//
// C-ABI wrapper around the UNMODIFIED reference sources (compiled where they lie under /root/reference by
// oracle/Makefile against the GLM-subset shim in oracle/shim/).  It is the "real reference" leg of the oracle:
//   * ref_load           = App::Run's scene set-up, /root/reference/src/App.cpp:115-130
//   * ref_intersect      = scene.hittable_list.Hit(...) exactly as RayColor issues it, RayTracer.cpp:25
//   * ref_render         = the per-pixel lambda of RayTracer::Update, RayTracer.cpp:55-70, with our own std::thread
//                          row striping (TBB is not installed, so std::execution::par would be serial) and per-sample
//                          first/second moments in double for the z-score image test
//   * ref_tracer_*       = the real RayTracer object (Update/Reset/OnResize/Pixels/NonConvertedPixels), serial
//   * ref_perlin_get/set = read / overwrite the per-load random Perlin tables (PerlinNoiseGen.cpp:41-50) so the GPU
//                          side and the oracle evaluate the same noise field
//   * ref_bvh_span1      = which top-level objects sit in a span-1 BVH leaf (BVH.cpp:18-20) -> Q2 double sampling
// The reference's arithmetic is not altered anywhere; private members are reached with `#define private public`
// in THIS translation unit only (no layout change).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>

#define private public
#include "cpu_raytrace/BVH.hpp"
#include "cpu_raytrace/ConstantMedium.hpp"
#include "cpu_raytrace/PerlinNoiseGen.hpp"
#include "cpu_raytrace/Texture.hpp"
#undef private

#include "Serialize.hpp"
#include "Util.hpp"
#include "cpu_raytrace/Scene.hpp"
// Pulls RayColor (anonymous namespace) and the RayTracer member definitions into this TU; RayTracer.cpp is therefore
// NOT compiled separately by the Makefile.
#include "cpu_raytrace/RayTracer.cpp"

using namespace raytrace2;

namespace {

thread_local uint64_t tl_ray_count = 0;

// Counts closest-hit queries (= "rays") without touching reference code: RayColor calls
// scene.hittable_list.Hit -> this proxy -> the BVH root.
struct CountingProxy : public cpu::Hittable {
  explicit CountingProxy(std::shared_ptr<cpu::Hittable> inner) : inner_(std::move(inner)) {}
  bool Hit(const cpu::Scene& scene, const cpu::Ray& r, cpu::Interval ray_t, cpu::HitRecord& rec) const override {
    tl_ray_count++;
    return inner_->Hit(scene, r, ray_t, rec);
  }
  [[nodiscard]] cpu::AABB GetAABB() const override { return inner_->GetAABB(); }
  std::shared_ptr<cpu::Hittable> inner_;
};

struct RefScene {
  cpu::Scene scene;
  std::vector<std::shared_ptr<cpu::Hittable>> top_level;  // scene array order, before the BVH constructor sorts
  std::shared_ptr<cpu::BVHNode> bvh;
  int w{0}, h{0};
  std::unique_ptr<cpu::RayTracer> tracer;
};

cpu::PerlinNoiseGen* NoiseGenAt(RefScene* s, int tex_idx) {
  if (tex_idx < 0 || tex_idx >= static_cast<int>(s->scene.textures.size())) return nullptr;
  auto* n = std::get_if<cpu::texture::Noise>(&s->scene.textures[tex_idx]);
  return n ? &n->noise : nullptr;
}

void CollectSpan1(const cpu::BVHNode* node, const RefScene* s, uint8_t* flags) {
  if (node->left_ == node->right_) {
    for (size_t i = 0; i < s->top_level.size(); i++)
      if (s->top_level[i] == node->left_) flags[i] = 1;
    return;
  }
  for (const auto& child : {node->left_, node->right_}) {
    if (const auto* inner = dynamic_cast<const cpu::BVHNode*>(child.get())) CollectSpan1(inner, s, flags);
  }
}

}  // namespace

extern "C" {

// App.cpp:115-130: load, apply scene dims (else 1600x900), wrap the top-level list in a BVH, set spp on the camera.
void* ref_load(const char* path, int num_samples, int width_override, int height_override) {
  serialize::SceneLoader loader;
  std::optional<cpu::Scene> opt;
  try {
    opt = loader.LoadScene(path);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "ref_load: reference loader threw: %s\n", e.what());
    return nullptr;
  }
  if (!opt.has_value()) return nullptr;
  auto* s = new RefScene;
  s->scene = std::move(opt.value());
  glm::ivec2 dims{1600, 900};
  if (s->scene.dims.x != 0 && s->scene.dims.y != 0) dims = s->scene.dims;
  if (width_override > 0 && height_override > 0) dims = {width_override, height_override};
  s->scene.cam.SetDims(dims);
  s->w = dims.x;
  s->h = dims.y;
  s->top_level = s->scene.hittable_list.objects;
  s->bvh = std::make_shared<cpu::BVHNode>(s->scene.hittable_list);
  s->scene.hittable_list = cpu::HittableList{std::make_shared<CountingProxy>(s->bvh)};
  s->scene.cam.SetSamplesPerPixel(num_samples);
  s->scene.cam.Update();
  return s;
}

void ref_free(void* h) { delete static_cast<RefScene*>(h); }

void ref_info(void* h, int* w, int* hgt, int* n_materials, int* n_textures, int* n_top, float* bg) {
  auto* s = static_cast<RefScene*>(h);
  *w = s->w;
  *hgt = s->h;
  *n_materials = static_cast<int>(s->scene.materials.size());
  *n_textures = static_cast<int>(s->scene.textures.size());
  *n_top = static_cast<int>(s->top_level.size());
  for (int i = 0; i < 3; i++) bg[i] = s->scene.background_color[i];
}

// out[0..2]=center, [3..5]=pixel00, [6..8]=delta_u, [9..11]=delta_v, [12..14]=defocus_u, [15..17]=defocus_v,
// [18]=defocus_angle, [19]=sqrt_spp
void ref_camera(void* h, float* out) {
  auto* s = static_cast<RefScene*>(h);
  const cpu::Camera& c = s->scene.cam;
  const vec3* v[6] = {&c.center_, &c.pixel00_loc_, &c.pixel_delta_u_, &c.pixel_delta_v_, &c.defocus_disk_u_, &c.defocus_disk_v_};
  for (int i = 0; i < 6; i++)
    for (int k = 0; k < 3; k++) out[i * 3 + k] = (*v[i])[k];
  out[18] = c.defocus_angle_;
  out[19] = static_cast<float>(c.SqrtSamplesPerPixel());
}

// rays: n x 7 floats (origin, direction, time).  Closest hit over [tmin, tmax] exactly as RayTracer.cpp:25 issues it.
// mat = index into scene.materials (or -1).  Stochastic for scenes holding constant media.
void ref_intersect(void* h, const float* rays, size_t n, float tmin, float tmax, uint8_t* hit, float* t, float* point,
                   float* normal, uint8_t* front_face, int32_t* mat) {
  auto* s = static_cast<RefScene*>(h);
  const cpu::Scene& scene = s->scene;
  for (size_t i = 0; i < n; i++) {
    const float* r = rays + i * 7;
    cpu::Ray ray{.origin = vec3{r[0], r[1], r[2]}, .direction = vec3{r[3], r[4], r[5]}, .time = r[6]};
    cpu::HitRecord rec;
    bool got = scene.hittable_list.Hit(scene, ray, cpu::Interval{tmin, tmax}, rec);
    hit[i] = got ? 1 : 0;
    if (got) {
      t[i] = rec.t;
      for (int k = 0; k < 3; k++) {
        point[i * 3 + k] = rec.point[k];
        normal[i * 3 + k] = rec.normal[k];
      }
      front_face[i] = rec.front_face ? 1 : 0;
      mat[i] = static_cast<int32_t>(rec.material - scene.materials.data());
    } else {
      t[i] = 0;
      for (int k = 0; k < 3; k++) point[i * 3 + k] = normal[i * 3 + k] = 0;
      front_face[i] = 0;
      mat[i] = -1;
    }
  }
}

// The per-pixel body of RayTracer::Update (RayTracer.cpp:57-67) for frames [frame0, frame0+nframes), rows striped over
// nthreads std::threads.  sum / sumsq: W*H*3 doubles (row 0 = bottom of the image, like accumulation_data_), ADDED to.
void ref_render(void* h, int frame0, int nframes, int max_depth, int nthreads, double* sum, double* sumsq,
                uint64_t* n_rays, double* seconds) {
  auto* s = static_cast<RefScene*>(h);
  const cpu::Scene& scene = s->scene;
  const cpu::Camera& cam = s->scene.cam;
  const int W = s->w, H = s->h;
  const int sq = cam.SqrtSamplesPerPixel();
  if (nthreads < 1) nthreads = 1;
  std::vector<uint64_t> counts(nthreads, 0);
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int tid = 0; tid < nthreads; tid++) {
    pool.emplace_back([&, tid]() {
      tl_ray_count = 0;
      for (int f = frame0; f < frame0 + nframes; f++) {
        int s_i = f % sq;
        int s_j = f / sq % sq;
        for (int y = tid; y < H; y += nthreads) {
          for (int x = 0; x < W; x++) {
            vec3 c = cpu::RayColor(cam.GetRay(x, y, s_i, s_j), max_depth, scene);
            size_t idx = (static_cast<size_t>(y) * W + x) * 3;
            for (int k = 0; k < 3; k++) {
              double v = c[k];
              sum[idx + k] += v;
              if (sumsq) sumsq[idx + k] += v * v;
            }
          }
        }
      }
      counts[tid] = tl_ray_count;
    });
  }
  for (auto& th : pool) th.join();
  auto t1 = std::chrono::steady_clock::now();
  uint64_t total = 0;
  for (auto c : counts) total += c;
  if (n_rays) *n_rays = total;
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
}

// ---- the real RayTracer object, serial (RayTracer.hpp:15-42) ----
void ref_tracer_init(void* h, int max_depth) {
  auto* s = static_cast<RefScene*>(h);
  s->tracer = std::make_unique<cpu::RayTracer>();
  s->tracer->max_depth = max_depth;
  s->tracer->camera = &s->scene.cam;
  s->tracer->OnResize(glm::ivec2{s->w, s->h});
}
void ref_tracer_update(void* h, int n) {
  auto* s = static_cast<RefScene*>(h);
  for (int i = 0; i < n; i++) s->tracer->Update(s->scene);
}
void ref_tracer_reset(void* h) { static_cast<RefScene*>(h)->tracer->Reset(); }
uint64_t ref_tracer_frame_idx(void* h) { return static_cast<RefScene*>(h)->tracer->FrameIdx(); }
void ref_tracer_read(void* h, float* mean_rgb, uint8_t* rgba8) {
  auto* s = static_cast<RefScene*>(h);
  if (mean_rgb) {
    auto px = s->tracer->NonConvertedPixels();
    std::memcpy(mean_rgb, px.data(), px.size() * sizeof(vec3));
  }
  if (rgba8) {
    const auto& px = s->tracer->Pixels();
    std::memcpy(rgba8, px.data(), px.size() * 4);
  }
}
// util::WriteImage (Util.cpp:39-79) on caller-supplied float RGB (row 0 = bottom).
void ref_write_image(const float* rgb, int w, int h, const char* path, int png) {
  std::vector<vec3> px(static_cast<size_t>(w) * h);
  std::memcpy(px.data(), rgb, px.size() * sizeof(vec3));
  util::WriteImage(px, w, h, path, png != 0);
}

// ---- Perlin tables (PerlinNoiseGen.hpp:15-20) ----
int ref_perlin_point_count(void* h, int tex_idx) {
  auto* g = NoiseGenAt(static_cast<RefScene*>(h), tex_idx);
  return g ? g->point_count_ : -1;
}
int ref_perlin_get(void* h, int tex_idx, int32_t* px, int32_t* py, int32_t* pz, float* vecs) {
  auto* g = NoiseGenAt(static_cast<RefScene*>(h), tex_idx);
  if (!g) return -1;
  for (int i = 0; i < g->point_count_; i++) {
    px[i] = g->perm_x_[i];
    py[i] = g->perm_y_[i];
    pz[i] = g->perm_z_[i];
    for (int k = 0; k < 3; k++) vecs[i * 3 + k] = g->rand_vec3_[i][k];
  }
  return g->point_count_;
}
int ref_perlin_set(void* h, int tex_idx, const int32_t* px, const int32_t* py, const int32_t* pz, const float* vecs) {
  auto* g = NoiseGenAt(static_cast<RefScene*>(h), tex_idx);
  if (!g) return -1;
  for (int i = 0; i < g->point_count_; i++) {
    g->perm_x_[i] = px[i];
    g->perm_y_[i] = py[i];
    g->perm_z_[i] = pz[i];
    g->rand_vec3_[i] = vec3{vecs[i * 3], vecs[i * 3 + 1], vecs[i * 3 + 2]};
  }
  return g->point_count_;
}
// Evaluate texture tex_idx at n points (deterministic once the tables are fixed): Texture.cpp:7-22.
void ref_texture_value(void* h, int tex_idx, const float* pts, size_t n, float* rgb) {
  auto* s = static_cast<RefScene*>(h);
  for (size_t i = 0; i < n; i++) {
    vec3 p{pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]};
    vec3 c = std::visit([&](auto&& tex) -> vec3 { return tex.Value(s->scene.textures, vec2{0, 0}, p); },
                        s->scene.textures[tex_idx]);
    for (int k = 0; k < 3; k++) rgb[i * 3 + k] = c[k];
  }
}

// flags[i] = 1 iff top-level scene object i was placed in a span-1 leaf (left_ == right_) by BVH.cpp:18-20.
int ref_bvh_span1(void* h, uint8_t* flags) {
  auto* s = static_cast<RefScene*>(h);
  std::memset(flags, 0, s->top_level.size());
  CollectSpan1(s->bvh.get(), s, flags);
  return static_cast<int>(s->top_level.size());
}
// For each top-level object: 1 if it is (or directly wraps) a ConstantMedium.  Used with ref_bvh_span1.
int ref_top_level_is_medium(void* h, uint8_t* flags) {
  auto* s = static_cast<RefScene*>(h);
  for (size_t i = 0; i < s->top_level.size(); i++) {
    const cpu::Hittable* o = s->top_level[i].get();
    flags[i] = dynamic_cast<const cpu::ConstantMedium*>(o) != nullptr;
  }
  return static_cast<int>(s->top_level.size());
}

}  // extern "C"
