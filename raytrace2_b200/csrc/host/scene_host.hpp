// Host scene compiler: JSON scene (current + legacy format) -> scene graph -> flat SoA buffers + BVH.
// Replaces serialize::SceneLoader::LoadScene (src/Serialize.cpp:199-360) and the BVH wrap of App::Run
// (src/App.cpp:126).  See scene_host.cpp for the per-step reference citations.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/rt2.h"
#include "hostmath.hpp"

namespace rt2 {

// Camera parameters as the loader sets them (Serialize.cpp:32-40, Camera.hpp:113-131).
struct CameraParams {
  V3 center{0, 0, 0};
  V3 look_at{0, 0, -1};
  V3 view_up{0, 1, 0};
  float vfov{90.f};
  float defocus_angle{0};
  float focus_dist{10};
};

// BVH build input: one (box, primitive reference) record per leaf primitive of a tree.
struct BuildPrim {
  float bmin[3];
  float bmax[3];
  uint32_t ref;
};

struct HostScene {
  std::vector<rt2_sphere> spheres;
  std::vector<rt2_quad> quads;
  std::vector<rt2_xform> xforms;
  std::vector<rt2_instance> instances;
  std::vector<rt2_medium> media;
  std::vector<float> media_bounds;  // 8 floats per medium: padded AABB of its boundary in the medium's own space (min xyz 0, max xyz 0)
  std::vector<rt2_material> materials;
  std::vector<rt2_texture> textures;
  std::vector<rt2_perlin> perlin;
  std::vector<rt2_image> images;     // image textures (schema extension): table + one shared texel array
  std::vector<float> image_texels;   // RGBA float32, linear light, row 0 = top
  std::vector<uint32_t> prim_refs;
  std::vector<rt2_bvh_node> nodes;  // 2 per pair
  uint32_t tlas_root{0};
  // Instance split (device/rt_trace.cuh kTravWorld): scenes with 1..RT2_MAX_HOISTED_INSTANCES instances also get a world TLAS
  // over the surfaces only; the instances are then tested by their world boxes (inst_bounds) after the world walk.
  uint32_t tlas_world_root{0};
  bool has_world_tlas{false};
  std::vector<float> inst_bounds;  // 8 floats per instance: conservative world-space AABB (min xyz 0, max xyz 0)
  // Unified world tree (device/rt_trace.cuh kTravUnified): every primitive of every instance is ALSO a leaf of one world-space
  // tree, with a conservative world-space box; leaf reference (RT2_PRIM_INSTANCE << 28 | k) names inst_leaves[k] =
  // {primitive reference, instance index}.  The primitive is still tested in the instance's model space, so hits are the
  // ones the two-level walk reports — only the culling structure is flat.  Built when the instances hold <= kMaxUnifiedLeaves
  // primitives in total.
  std::vector<uint32_t> inst_leaves;  // 2 per instanced leaf
  uint32_t tlas_unified_root{0};
  bool has_unified_tlas{false};
  std::vector<BuildPrim> unified_prims;  // build input of the unified tree (device LBVH build)
  uint32_t n_top_level{0};
  std::vector<uint8_t> span1_flags;  // per top-level node (Q2)
  // Leaf records of every tree, kept for the device-side LBVH build (RT2_FLAG_GPU_LBVH): [0] = world TLAS,
  // [1 + i] = BLAS of instance i.  has_host_bvh is false when the host SAH build was skipped (huge synthetic scenes).
  std::vector<std::vector<BuildPrim>> tree_prims;
  bool has_host_bvh{true};
  float background[3]{1, 1, 1};
  float min_inv_scale{1.f};
  int width{1600}, height{900};
  CameraParams cam;
  rt2_camera camera_block{};
  std::vector<std::string> warnings;

  void FillDesc(rt2_scene_desc* d) const;
  void UpdateCamera();  // Camera::Update, Camera.hpp:16-48
};

// Returns RT2_OK or a negative RT2_ERR_*; `err` gets a one-line reason.
int LoadSceneFile(const std::string& path, const std::string& data_dir, uint64_t perlin_seed, HostScene* out,
                  std::string* err);
int LoadSceneString(const std::string& text, const std::string& data_dir, uint64_t perlin_seed, HostScene* out,
                    std::string* err);
int MakeSyntheticSpheres(uint32_t n, uint64_t seed, int width, int height, bool build_host_bvh, HostScene* out, std::string* err);

// AppSettings (src/Settings.hpp:5-11, Serialize.cpp:56-65)
struct AppSettings {
  bool render_once{false};
  bool save_after_render_once{false};
  size_t num_samples{1};
  size_t max_depth{50};
  bool render_window{true};
};
int LoadAppSettings(const std::string& path, AppSettings* out, std::string* err);

// BVH builder over (box, ref) leaves; appends node pairs to scene.nodes and refs to scene.prim_refs; returns the
// root pair index.
uint32_t BuildBVH(std::vector<BuildPrim>& prims, HostScene* scene);

}  // namespace rt2
