/* include/rt2.h — C ABI of the B200-native path-tracing backend (libraytrace2_b200.so).
 *
 * The reference (tonadr1022/Raytrace2) has NO plugin / FFI layer: its renderer seam is a C++ struct used directly by
 * App.  This header is the drop-in boundary a maintainer binds instead of `raytrace2::cpu::RayTracer`,
 * `serialize::SceneLoader` and `util::WriteImage`; every entry point cites the reference interface it replaces
 * (paths relative to the reference root).  Plain C types only: pointers + sizes, caller owns all host memory,
 * every call returns 0 on success or a negative RT2_ERR_* code (message via rt2_last_error()); nothing throws across
 * the boundary.  One host thread per handle.  There is no CPU fallback: rt2_create fails if no CUDA device is usable.
 */
#ifndef RT2_H_
#define RT2_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT2_ABI_VERSION 3

enum {
  RT2_OK = 0,
  RT2_ERR_INVALID_ARG = -1,
  RT2_ERR_IO = -2,       /* file missing / unreadable */
  RT2_ERR_PARSE = -3,    /* malformed JSON or scene semantics the loader rejects (Serialize.cpp:246-249) */
  RT2_ERR_CUDA = -4,     /* any CUDA runtime failure, including "no device" */
  RT2_ERR_UNSUPPORTED = -5,
  RT2_ERR_STATE = -6
};

/* ---- flat (SoA) scene layout shared by host and device -------------------------------------------------------- */

/* primitive reference stored in BVH leaves and hit records: (type << 28) | index */
#define RT2_PRIM_SPHERE 0u
#define RT2_PRIM_QUAD 1u
#define RT2_PRIM_INSTANCE 2u
#define RT2_PRIM_MEDIUM 3u
#define RT2_PRIM_TYPE(ref) ((ref) >> 28)
#define RT2_PRIM_INDEX(ref) ((ref) & 0x0FFFFFFFu)
#define RT2_PRIM_NONE 0xFFFFFFFFu

/* material types (Material.hpp:31-65); RT2_MAT_TEXTURE is the reference's "MaterialTexture" (textured lambertian) */
enum { RT2_MAT_LAMBERTIAN = 0, RT2_MAT_METAL = 1, RT2_MAT_DIELECTRIC = 2, RT2_MAT_TEXTURE = 3, RT2_MAT_DIFFUSE_LIGHT = 4,
       RT2_MAT_ISOTROPIC = 5, RT2_MAT_INVALID = 6 };
/* texture types (Texture.hpp:14-36).  RT2_TEX_IMAGE is a schema extension ({"type": "image", "path": ...}):
 * the reference carries (u, v) through Texture::Value (Texture.hpp:15, Sphere.cpp:34,39-43, Quad.cpp:15) but has no texture
 * that reads them; the lookup follows the book the reference implements (nearest texel, v flipped, clamped). */
enum { RT2_TEX_SOLID = 0, RT2_TEX_CHECKER = 1, RT2_TEX_NOISE = 2, RT2_TEX_IMAGE = 3, RT2_TEX_INVALID = 4 };

/* Sphere.hpp:14-34 — centre(time) = center0 + time * displacement */
typedef struct rt2_sphere {
  float center0[3];
  float radius;
  float displacement[3];
  uint32_t material;
} rt2_sphere; /* 32 B */

/* Quad.hpp:13-32 — normal, d, w are the constructor's values, bit for bit */
typedef struct rt2_quad {
  float normal[3];
  float d;
  float q[3];
  uint32_t material;
  float u[3];
  float pad0;
  float v[3];
  float pad1;
  float w[3];
  float pad2;
} rt2_quad; /* 80 B */

/* One level of an instance chain (TransformedHittable, Transform.hpp:20-36): affine 3x4, row-major rows of the
 * column-major GLM matrices: inv[r][c] = inv_model[c][r].  normal matrix = transpose(inverse(model)) = inv columns. */
typedef struct rt2_xform {
  float inv[3][4];   /* inverse model, rows */
  float model[3][4]; /* model, rows */
} rt2_xform;         /* 96 B */

/* A flattened TransformedHittable: the levels [chain_first, chain_first+chain_len) are applied outermost first, each
 * followed by a direction normalisation (Transform.cpp:13-20); blas_root is the node-pair index of its BVH. */
typedef struct rt2_instance {
  uint32_t chain_first;
  uint32_t chain_len;
  uint32_t blas_root;
  uint32_t top_level_node; /* index of the scene[] entry this instance came from */
} rt2_instance;            /* 16 B */

/* ConstantMedium.hpp:5-16.  The boundary is the closed list of primitives prim_refs[boundary_first .. +count). */
typedef struct rt2_medium {
  float neg_inv_density;
  uint32_t material;
  uint32_t boundary_first;
  uint32_t boundary_count;
  uint32_t chain_first; /* instance chain the medium sits under (chain_len == 0: world space) */
  uint32_t chain_len;
  uint32_t sample_twice; /* 1 iff the owning top-level object is in a span-1 leaf of the reference's BVH (BVH.cpp:18-20) */
  uint32_t top_level_node;
} rt2_medium; /* 32 B */

typedef struct rt2_material {
  uint32_t type;
  uint32_t tex_idx;
  float fuzz;
  float refraction_index;
  float albedo[3];
  float pad;
} rt2_material; /* 32 B */

typedef struct rt2_texture {
  uint32_t type;
  uint32_t even_tex_idx; /* checker */
  uint32_t odd_tex_idx;  /* checker */
  uint32_t perlin_idx;   /* noise: index into the perlin table array */
  float albedo[3];       /* solid colour / noise albedo */
  float scale;           /* checker: inv_scale (= 1.f / scale, Texture.hpp:21); noise: scale */
  uint32_t noise_type;   /* 0 = perlin, 1 = marble (Texture.hpp:30) */
  uint32_t image_idx;    /* image: index into the image table */
  uint32_t pad[2];
} rt2_texture; /* 48 B */

/* One decoded image: texels are linear-light RGBA float32, row 0 = top of the picture, stored in one shared array. */
typedef struct rt2_image {
  uint32_t texel_offset; /* first texel (in float4 units) */
  uint32_t width, height;
  uint32_t pad;
} rt2_image; /* 16 B */

/* PerlinNoiseGen.hpp:15-20; the hash mask is 255 regardless of point_count (PerlinNoiseGen.cpp:83), so tables are
 * stored with 256 entries (entries >= point_count are never produced by a valid permutation and are zero). */
typedef struct rt2_perlin {
  int32_t perm_x[256];
  int32_t perm_y[256];
  int32_t perm_z[256];
  float vec[256][4]; /* xyz + pad */
} rt2_perlin;

/* Instance split: a scene with 1..RT2_MAX_HOISTED_INSTANCES instances is traversed in two passes — the world-space surfaces
 * first (a TLAS without instance leaves), then one {ray, instance} entry per instance whose world box the ray's remaining
 * segment touches, with the world-to-model transforms converged across the warp (device/rt_trace.cuh). */
#define RT2_MAX_HOISTED_INSTANCES 4

/* BVH node, 32 bytes; nodes are stored as sibling PAIRS (64-byte aligned): pair p = nodes[2p], nodes[2p+1].
 * count == 0: interior, left_first = index of the child pair.  count > 0: leaf over prim_refs[left_first .. +count).
 * An empty slot has NaN bounds (never entered by the slab test). */
typedef struct rt2_bvh_node {
  float bmin[3];
  uint32_t left_first;
  float bmax[3];
  uint32_t count;
} rt2_bvh_node;

/* Camera.hpp:113-131 after Camera::Update() */
typedef struct rt2_camera {
  float center[3];
  float pixel00[3];
  float pixel_delta_u[3];
  float pixel_delta_v[3];
  float defocus_disk_u[3];
  float defocus_disk_v[3];
  float defocus_angle;
  float vfov;
  float focus_dist;
  float look_at[3];
} rt2_camera;

typedef struct rt2_scene_desc {
  uint32_t n_spheres, n_quads, n_xforms, n_instances, n_media, n_materials, n_textures, n_perlin;
  uint32_t n_prim_refs, n_node_pairs, tlas_root; /* tlas_root: node-pair index of the world-space BVH */
  uint32_t n_top_level;                          /* entries of the JSON "scene" array */
  const rt2_sphere* spheres;
  const rt2_quad* quads;
  const rt2_xform* xforms;
  const rt2_instance* instances;
  const rt2_medium* media;
  const rt2_material* materials;
  const rt2_texture* textures;
  const rt2_perlin* perlin;
  const uint32_t* prim_refs;
  const rt2_bvh_node* nodes; /* 2 * n_node_pairs entries */
  float background[3];
  float min_inv_scale; /* smallest singular value over all instance chains' inverse 3x3 (1 for rigid chains) */
  int32_t width, height; /* scene-authored dims, or 1600x900 (App.cpp:115,122-125) */
  rt2_camera camera;     /* computed for (width, height) */
  uint32_t n_images;
  uint32_t n_image_texels;   /* float4 texels over all images */
  const rt2_image* images;
  const float* image_texels; /* 4 * n_image_texels floats */
  /* unified world tree: the world surfaces + every instanced primitive as a world-space leaf (leaf reference
   * RT2_PRIM_INSTANCE << 28 | k names inst_leaves[k] = {primitive reference, instance index}) */
  uint32_t has_unified_tlas;
  uint32_t tlas_unified_root;
  uint32_t n_inst_leaves;
  uint32_t pad_unified;
  const uint32_t* inst_leaves; /* 2 * n_inst_leaves */
  /* instance split (RT2_MAX_HOISTED_INSTANCES): a second world tree over the surfaces only, and the world boxes of the instances */
  uint32_t has_world_tlas;   /* 1 iff tlas_world_root is valid (the scene has 1..RT2_MAX_HOISTED_INSTANCES instances) */
  uint32_t tlas_world_root;  /* node-pair index */
  const float* inst_bounds;  /* 8 floats per instance: min xyz, 0, max xyz, 0 (conservative, world space) */
} rt2_scene_desc;

/* ---- scene: replaces serialize::SceneLoader::LoadScene (src/Serialize.hpp:21-22, Serialize.cpp:199-360) plus the
 * dims / BVH set-up of App::Run (src/App.cpp:115-126) --------------------------------------------------------------- */
typedef struct rt2_scene rt2_scene;

/* data_dir: directory holding named camera files ("camera": "<name>" -> <data_dir>/<name>.json, Serialize.cpp:208-209)
 * and the default camera for legacy scenes without one; NULL = directory of json_path.
 * perlin_seed: seeds the host generator of the random Perlin tables (the reference seeds from random_device). */
int rt2_scene_load(const char* json_path, const char* data_dir, uint64_t perlin_seed, rt2_scene** out);
int rt2_scene_load_string(const char* json_text, const char* data_dir, uint64_t perlin_seed, rt2_scene** out);
/* Synthetic BVH stress scene: n random spheres + ground sphere, book-1 material recipe.
 * build_host_bvh = 0 skips the host SAH build (seconds per million spheres); such a scene can only be rendered with
 * RT2_FLAG_GPU_LBVH. */
int rt2_scene_synthetic_spheres(uint32_t n_spheres, uint64_t seed, int32_t width, int32_t height, int32_t build_host_bvh,
                                rt2_scene** out);
void rt2_scene_destroy(rt2_scene* scene);
/* Borrowed pointers into the scene, valid until the scene is modified or destroyed. */
int rt2_scene_get_desc(const rt2_scene* scene, rt2_scene_desc* out);
/* Recompute the camera block for new dims (Camera::SetDims + Update, Camera.hpp:16-48,89-92). */
int rt2_scene_set_dims(rt2_scene* scene, int32_t width, int32_t height);
/* Overwrite / read the Perlin tables of perlin slot `perlin_idx` (test hook: share one noise field with the oracle).
 * perm_*: 256 int32 each, vec: 256*3 floats. */
int rt2_scene_set_perlin(rt2_scene* scene, uint32_t perlin_idx, const int32_t* perm_x, const int32_t* perm_y,
                         const int32_t* perm_z, const float* vec);
int rt2_scene_get_perlin(const rt2_scene* scene, uint32_t perlin_idx, int32_t* perm_x, int32_t* perm_y, int32_t* perm_z,
                         float* vec);
/* Per top-level scene[] entry: 1 iff it lands in a span-1 leaf of the reference-order BVH (Q2). flags: n_top_level bytes. */
int rt2_scene_span1_flags(const rt2_scene* scene, uint8_t* flags);

/* ---- renderer: replaces raytrace2::cpu::RayTracer (src/cpu_raytrace/RayTracer.hpp:15-42) ------------------------ */
typedef struct rt2_renderer rt2_renderer;

#define RT2_FLAG_MOMENTS 1u     /* also accumulate per-pixel sum of squares (z-score parity test) */
#define RT2_FLAG_FAST_MATH 2u   /* FMA-contracted intersection arithmetic (default: bit-exact with the reference) */
#define RT2_FLAG_NO_FUSED_SHADE 4u /* run k_finish_hit + one shade kernel per material bin instead of the fused
                                     k_finish_shade (A/B and debugging; results are identical) */
#define RT2_FLAG_GPU_LBVH 8u    /* build every BVH on the device (63-bit Morton codes + radix sort + Karras radix tree) instead of
                                   uploading the host SAH trees */
#define RT2_FLAG_WIDE_BVH 32u   /* with RT2_FLAG_GPU_LBVH, scenes without instances: collapse the device-built tree into 4-wide
                                   nodes with 8-bit child boxes (64 B per node) and walk those — half the node traffic of the
                                   binary tree; for scenes whose tree does not fit the caches (1 M - 10 M primitives) */
#define RT2_FLAG_SORT_RAYS 16u  /* reorder every bounce's ray queue by (direction octant, origin cell) before the traversal
                                   (device radix sort); changes the traversal ORDER only, never a result.  Measured as a net
                                   loss: compiled only into `make EXPERIMENTS=1` builds, RT2_ERR_UNSUPPORTED otherwise */
#define RT2_FLAG_NO_FLAT_EXTEND 128u /* walk the BVH even in tiny scenes that would take the tree-less flat extend kernel (A/B) */
/* How instances are walked.  Default: ONE world-space tree whose leaves are the world surfaces and every primitive of every
 * instance (each still tested in its instance's model space) — scenes whose instances hold <= 4 M primitives in total. */
#define RT2_FLAG_INSTANCES_INLINE 64u /* two-level walk in one kernel: instance leaves in the TLAS, a lane that meets one switches
                                   to the model space and the BLAS (always used by scenes too big to flatten); A/B and debugging */
#define RT2_FLAG_NO_INSTANCE_SPLIT RT2_FLAG_INSTANCES_INLINE /* (round-2 name) */
#define RT2_FLAG_LBVH_PLOC 512u /* with RT2_FLAG_GPU_LBVH: build the hierarchy by PLOC (parallel locally-ordered clustering over the
                                   Morton order) instead of Karras' one-pass radix tree.  Measured WORSE on the sphere stress scene
                                   (38.7 vs 30.0 node pairs per ray at 1 M spheres): opt-in, for A/B */
#define RT2_FLAG_FLOAT_NODES 1024u /* keep the 64-byte float node pairs even where the 32-byte quantised pairs would be used (A/B;
                                     same hits: box tests only cull) */
#define RT2_FLAG_INSTANCE_SPLIT 256u /* two passes: surfaces-only world tree, then one {ray, instance} entry per touched instance
                                   (RT2_MAX_HOISTED_INSTANCES); measured equal to the inline walk, kept for A/B */

typedef struct rt2_config {
  int32_t device;            /* CUDA device ordinal (the first one when n_gpus > 1) */
  int32_t width, height;     /* 0 = scene dims */
  int32_t samples_per_pixel; /* AppSettings::num_samples -> Camera::SetSamplesPerPixel (App.cpp:129): stratification grid */
  int32_t max_depth;         /* AppSettings::max_depth -> RayTracer::max_depth (App.cpp:128) */
  int32_t frames_per_batch;  /* frames (samples per pixel) traced per wavefront batch; 0 = auto */
  int32_t frame_offset;      /* multi-GPU sample partition: this renderer traces global frames */
  int32_t frame_stride;      /*   frame_offset + k * frame_stride, k = 0,1,...  (stride 0 or 1 = all frames) */
  uint32_t flags;
  uint64_t seed;             /* Philox key; the reference seeds from std::random_device (Math.hpp:11) */
  int32_t n_gpus;            /* 0 or 1: one GPU (`device`).  N > 1: devices device .. device+N-1 of this box.  -1: every visible
                                device.  The handle then owns one replica per GPU in THIS process: the frames
                                of every rt2_update are dealt round-robin (global frame f goes to replica f mod N, which keeps
                                each GPU's strata spread over the sqrt(spp) x sqrt(spp) grid, RayTracer.cpp:59-60), and every
                                read-out sums the replicas' accumulators in replica order inside ONE kernel on the first GPU
                                that reads the peers' HBM over NVLink (cudaDeviceEnablePeerAccess; staged copies when two
                                devices have no peer path).  Images differ from the 1-GPU image by fp32 summation order only. */
  int32_t reserved;
} rt2_config;

typedef struct rt2_stats {
  uint64_t rays;        /* closest-hit queries issued since the last reset (RayTracer.cpp:25) */
  uint64_t paths;       /* camera paths started since the last reset */
  uint64_t frames;      /* local frames rendered since the last reset */
  uint64_t launches;    /* kernels launched since creation */
  double gpu_ms_total;  /* CUDA-event time of all rt2_update calls since the last reset */
  double gpu_ms_extend; /* summed CUDA-event time of k_traverse (only while profiling is enabled) */
  double gpu_ms_shade;  /* ... of the shade kernels */
  double gpu_ms_other;  /* ... of generate / accumulate / stats */
  double gpu_ms_finish; /* ... of k_finish_hit (media + hit record + binning) */
  double gpu_ms_bvh_build; /* CUDA-event time of the last device-side LBVH build (RT2_FLAG_GPU_LBVH), else 0 */
  /* algorithmic work done by active lanes of k_traverse while profiling is enabled */
  uint64_t box_pair_tests; /* node-pair visits = 2 AABB slab tests each */
  uint64_t sphere_tests;
  uint64_t quad_tests;
  uint64_t instance_visits;
  double gpu_ms_sort; /* ... of the ray-sort kernels (RT2_FLAG_SORT_RAYS) */
  uint64_t stack_overflows; /* warps whose traversal found its stack full (wide-BVH walk only; the binary walks cannot overflow:
                               rt2_create / rt2_upload_scene refuse a scene whose trees are deeper than the stack).  0 in a
                               valid render; read-outs fail otherwise */
  uint64_t pending_frames;  /* always 0 here: rt2_get_stats traces every requested frame first */
  uint32_t n_gpus;          /* replicas behind this handle */
  uint32_t instance_split;  /* 1 iff the two-pass instance split is active */
  double gpu_ms_extend_inst; /* instance split, while profiling: time of the instance pass (gpu_ms_extend = the world pass) */
  uint32_t max_stack_need;  /* stack entries the deepest traversal of this scene can need (tree depths, verified <= 63 at upload) */
  uint32_t instance_mode;   /* 0 no instances, 1 inline TLAS -> BLAS, 2 two-pass split, 3 unified world tree, 4 flat extend */
  uint32_t compact_nodes;   /* 1 iff the walk reads 32-byte node pairs with boxes quantised onto one 15-bit grid per scene */
  float node_inflation;     /* mean over the node boxes of (quantised / float surface area) - 1: the measure the choice is made on */
} rt2_stats;

typedef struct rt2_hit {
  float point[3];
  float t;
  float normal[3];
  int32_t material;    /* index into materials, -1 = miss */
  uint32_t prim;       /* RT2 prim ref of the closest leaf (RT2_PRIM_NONE on miss) */
  int32_t instance;    /* flattened instance index or -1 */
  uint32_t front_face; /* 0/1 */
  uint32_t uv16;       /* HitRecord::uv (Sphere.cpp:34, Quad.cpp:15) as two 16-bit unorms, u | v << 16; 0 for media and misses */
} rt2_hit; /* 48 B */

/* Uploads the flattened scene to `cfg->device` and allocates the wavefront state. ≡ RayTracer + OnResize. */
int rt2_create(const rt2_scene* scene, const rt2_config* cfg, rt2_renderer** out);
void rt2_destroy(rt2_renderer* r);
/* Re-upload the scene buffers from host memory (used by the end-to-end benchmark; also after rt2_scene_set_perlin). */
int rt2_upload_scene(rt2_renderer* r, const rt2_scene* scene);
/* RayTracer::OnResize (RayTracer.cpp:87-104): new dims on the camera, reallocate, Reset. */
int rt2_resize(rt2_renderer* r, int32_t width, int32_t height);
/* RayTracer::Reset (RayTracer.cpp:49-53): zero the accumulators and the frame index. */
int rt2_reset(rt2_renderer* r);
/* n_frames x RayTracer::Update (RayTracer.cpp:55-70): +1 sample per pixel each.  The reference's app calls Update once per
 * sample (App.cpp:244-246); here requested frames are collected until a wavefront batch (frames_per_batch per GPU) is full and
 * traced then, asynchronously on the renderer's stream(s) — or by the next call that needs them (any rt2_read_*, rt2_flush,
 * rt2_synchronize, rt2_get_stats, ...).  A frame's stratum and random numbers depend only on its index, so 10 000 calls with
 * n_frames = 1 give the same image as one call with n_frames = 10 000, at the same speed. */
int rt2_update(rt2_renderer* r, uint32_t n_frames);
/* Trace every frame requested so far (asynchronous); rt2_synchronize also waits for the GPU(s). */
int rt2_flush(rt2_renderer* r);
int rt2_synchronize(rt2_renderer* r);
/* RayTracer::FrameIdx (RayTracer.hpp:23) — frames requested from this renderer since the last reset (traced or pending). */
int rt2_frame_idx(const rt2_renderer* r, uint64_t* out);
/* RayTracer::Dims (RayTracer.hpp:29) */
int rt2_dims(const rt2_renderer* r, int32_t* width, int32_t* height);
/* RayTracer::NonConvertedPixels (RayTracer.cpp:105-112): mean = accum / frame_idx, W*H*3 floats, row 0 = bottom. */
int rt2_read_mean_rgb32f(rt2_renderer* r, float* dst);
/* RayTracer::Pixels (RayTracer.hpp:22, RayTracer.cpp:16-18,65-66): RGBA8, floor(255.999*clamp(mean,0,1)), no gamma. */
int rt2_read_rgba8(rt2_renderer* r, uint8_t* dst);
/* Raw accumulators: sum (W*H*3 floats) and, with RT2_FLAG_MOMENTS, sum of squares (else sumsq may be NULL). */
int rt2_read_accum(rt2_renderer* r, float* sum, float* sumsq);
/* Checkpoint / resume (the reference has none: a 10k-spp render that dies starts over): restore the accumulators read with
 * rt2_read_accum and the number of frames they hold.  Frames rendered afterwards continue the same stratification and Philox
 * counters, so an interrupted + resumed render is bit-identical with an uninterrupted one (same batch partition). */
int rt2_write_accum(rt2_renderer* r, const float* sum, const float* sumsq, uint64_t frames);
/* Device pointer of the W*H*4-float accumulator (rgb + pad) for an external reduce (NCCL / torch.distributed). */
int rt2_accum_device_ptr(rt2_renderer* r, void** ptr, size_t* n_floats);
/* Multi-GPU read-out over peer memory (one process per GPU on one NVLink / NVSwitch box).  Every rank exports
 * an RT2_IPC_HANDLE_BYTES-byte handle of its accumulator (CUDA IPC handle + offset inside the exported allocation); the reading
 * rank passes all ranks' handles (n_ranks x RT2_IPC_HANDLE_BYTES bytes, its own slot is ignored) and one kernel sums the peers' accumulators in rank order with P2P loads, divides by total_frames and writes
 * the mean (W*H*3 floats) and / or the RGBA8 preview — reduce + NonConvertedPixels / Pixels in one pass.  The CALLER
 * synchronises: every rank must have finished rendering (rt2_synchronize + a barrier) before the call, and must not
 * touch its accumulator until the reading rank returns.  Handles stay valid until the exporting renderer is resized or
 * destroyed. */
#define RT2_IPC_HANDLE_BYTES 80
int rt2_accum_ipc_handle(rt2_renderer* r, uint8_t* handle);
int rt2_resolve_peers(rt2_renderer* r, const uint8_t* handles, uint32_t n_ranks, uint32_t self_rank, uint64_t total_frames,
                      float* dst_mean_rgb, uint8_t* dst_rgba8);
/* After an external reduce: declare how many frames the accumulator now holds in total. */
int rt2_set_frame_idx(rt2_renderer* r, uint64_t frames);
/* Fixed-ray parity hook ≡ scene.hittable_list.Hit(ray, Interval{tmin,tmax}) (RayTracer.cpp:25). rays: n x 8 floats
 * (origin xyz, time, direction xyz, pad). Media are sampled with Philox keyed on the ray index unless skip_media != 0. */
int rt2_intersect(rt2_renderer* r, const float* rays, size_t n, float tmin, float tmax, int skip_media, rt2_hit* out);
/* Fixed-point parity hook for the device texture code ≡ textures[tex_idx]->Value(u, v, p) (Texture.hpp:14-17, Texture.cpp:7-22,
 * PerlinNoiseGen.cpp:52-88), evaluated by the same device function the shade kernels call.  points: n x 3 floats, uv: n x 2
 * floats or NULL (0, 0), rgb: n x 3 floats out. */
int rt2_texture_value(rt2_renderer* r, uint32_t tex_idx, const float* points, const float* uv, size_t n, float* rgb);
/* Copies the BVH the device traverses back to the host (inspection / tests): nodes = 2 * n_pairs entries. Pass NULL
 * buffers to query the sizes only. */
int rt2_read_bvh(rt2_renderer* r, rt2_bvh_node* nodes, size_t max_nodes, uint32_t* prim_refs, size_t max_refs, uint32_t* n_pairs,
                 uint32_t* n_refs, uint32_t* tlas_root);
int rt2_get_stats(rt2_renderer* r, rt2_stats* out);
/* Ray-queue sizes of the most recent wavefront batch on the first GPU: out[b] = rays traced at bounce b (b < max_bounces);
 * *n_bounces = max_depth.  For per-bounce analysis and for turning an ncu capture of one launch into bytes per ray. */
int rt2_read_queue_sizes(rt2_renderer* r, uint32_t* out, uint32_t max_bounces, uint32_t* n_bounces);
/* Device-side self checks.  compute-sanitizer is not available on every GPU pool, so a `make DEBUG_CHECKS=1` build of this
 * library (libraytrace2_b200_dbg.so) verifies every data-dependent index in the kernels (node, primitive, material, queue,
 * entry, stack ...) and counts violations per kind; it also poisons the wavefront buffers with NaN patterns at allocation.
 * counters: 16 x uint64 (first GPU of the handle); *enabled = 1 iff the library was built with the checks. */
int rt2_debug_counters(rt2_renderer* r, uint64_t* counters, int* enabled);
int rt2_set_profiling(rt2_renderer* r, int enabled); /* per-kernel CUDA-event timing (serialises launches) */
/* CUDA stream the renderer launches on (cudaStream_t as void*), for external event timing. */
int rt2_stream(rt2_renderer* r, void** stream);

/* ---- output: replaces util::WriteImage(vector<vec3>, w, h, path, png) (src/Util.hpp:11-12, Util.cpp:39-79) ------ */
int rt2_write_image(const float* mean_rgb, int32_t width, int32_t height, const char* out_path, int png);
/* The 8-bit conversion WriteImage applies (sqrt gamma, clamp(255.999*c, 0, 255)), top row first. dst: W*H*3 bytes. */
int rt2_tonemap_rgb8(const float* mean_rgb, int32_t width, int32_t height, uint8_t* dst);

/* ---- app: the headless branch of App::Run (src/App.cpp:81-130,157,163-174,243-248) -------------------------------
 * argv as given to `raytrace_2 [scene[.json]] [out.png]`; settings_path ≡ <SRC_PATH>/local/data/settings.json.  Renders on
 * every GPU of the box by default.  Extensions (removed from argv before the reference's positional parsing):
 * `--gpus N` (1..visible), `--spp N` (overrides settings.num_samples), `--max-depth N`, `--seed N`. */
int rt2_app_run(int argc, const char* const* argv, const char* settings_path, const char* data_dir);

const char* rt2_last_error(void);
int rt2_abi_version(void);
int rt2_device_count(void);
/* Measured FP32 FMA peak of a device in TFLOP/s (micro-benchmark kernel; the roofline denominator of the instruction-bound
 * kernels). */
int rt2_measure_fp32_peak(int32_t device, double* tflops);
/* Measured L2 read bandwidth of a device in GB/s (a working set that stays L2-resident, read repeatedly with 16-byte loads):
 * the roofline of traversals whose tree lives in L2 but not in L1. */
int rt2_measure_l2_bandwidth(int32_t device, double* gbs);

#ifdef __cplusplus
}
#endif
#endif /* RT2_H_ */
