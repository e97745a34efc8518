// Wavefront path-tracing kernels for sm_100a and their launcher.
//
// Replaces RayTracer::Update / RayColor (src/cpu_raytrace/RayTracer.cpp:20-70) with a batch-synchronous wavefront:
//
//   k_generate            Camera::GetRay for F frames x W*H pixels                       (Camera.hpp:50-67)
//   per bounce b < max_depth:
//     [k_sort_*]          optional: coherence order of the queue                          (RT2_FLAG_SORT_RAYS, rt_sort.cuh)
//     k_traverse          closest SURFACE: persistent while-while walk (rt_trace.cuh) of the unified world tree (scenes with
//                         instances) or the TLAS -> instance BLAS pair, over 32-byte quantised node pairs where the scene's
//                         leaves are large against the grid cell (rt_qnodes.cu), 64-byte float pairs otherwise;
//       | k_traverse_flat   tiny scenes: every primitive, uniform loops, no tree
//       | k_traverse_wide   RT2_FLAG_WIDE_BVH: 4-wide quantised nodes             (rt_wide.cuh)
//     [k_media]           scenes whose media have list boundaries: media sampling as its own pass
//     k_finish_shade      media (else), hit record, Material::Scatter / Emit inline, next ray appended to the next
//                         queue (one atomic per 256-ray tile, grouped by material class)  (RayTracer.cpp:25-44)
//     [k_shade_scatter<TEXTURE|ISOTROPIC>, k_shade_terminal]  only the noise-textured (deferred) materials' bins
//     (RT2_FLAG_NO_FUSED_SHADE: k_finish_hit -> hit records in HBM -> one k_shade_* launch per material bin)
//   k_accumulate          accum[pixel] += radiance[f][pixel] in frame order                (RayTracer.cpp:64)
//   k_resolve | k_resolve_peers   mean, RGBA8 preview; across GPUs over peer memory        (RayTracer.cpp:16-18,65-66,105-112)
//
// Queue sizes live in device memory (counters[bounce][kCounterStride]); kernels are launched with a persistent grid and loop
// grid-stride over the count they read there, so no host synchronisation happens inside a batch.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "rt_lbvh.hpp"
#include "rt_qnodes.hpp"
#include "rt_render.hpp"
#include "rt_shade.cuh"
#ifdef RT2_WITH_RAY_SORT
#include "rt_sort.cuh"  // measured net loss (profiles/r01_notes.md): only in `make EXPERIMENTS=1` builds
#endif
#include "rt_wide.cuh"
#include "../host/tuning.hpp"

namespace rt2dev {

constexpr int kBlock = 256;
constexpr int kNumBins = 6;  // 0 terminal (miss / light), 1 lambertian, 2 texture, 3 metal, 4 dielectric, 5 isotropic
constexpr int kCounterStride = 12;  // [0] queue size, [1..6] material bins, [7] traversal fetch cursor,
                                    // [8] instance-split entries of the bounce, [9] their fetch cursor
constexpr int kFetchThreshold = 20;  // lanes: below this a warp refills its idle lanes from the queue

struct FrameParams {
  rt2_camera cam;
  uint32_t width, height, pixels;
  uint32_t sqrt_spp;
  float recip_sqrt_spp;
  uint32_t frame_base;    // global index of the batch's first frame
  uint32_t frame_stride;  // global frame of local frame lf = frame_base + lf * frame_stride
  uint32_t seed_lo, seed_hi;
  unsigned long long pixels_magic;  // ceil(2^64 / pixels): slot / pixels == umul64hi(slot, magic) for slot < 2^32, pixels >= 2
};

__device__ __forceinline__ RngKey key_of_slot(const FrameParams& fp, uint32_t slot) {
  RngKey k;
  const uint32_t lf = fp.pixels > 1u ? static_cast<uint32_t>(__umul64hi(static_cast<unsigned long long>(slot), fp.pixels_magic)) : slot;
  k.pixel = slot - lf * fp.pixels;
  k.frame = fp.frame_base + lf * fp.frame_stride;
  k.seed_lo = fp.seed_lo;
  k.seed_hi = fp.seed_hi;
  return k;
}

// Camera::GetRay (Camera.hpp:50-67) + the stratum selection of RayTracer::Update (RayTracer.cpp:57-60)
__global__ void __launch_bounds__(kBlock) k_generate(const FrameParams fp, uint32_t n_slots, float4* __restrict__ ray_o,
                                                     float4* __restrict__ ray_d, float4* __restrict__ state,
                                                     uint32_t* __restrict__ counters) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_slots; slot += stride) {
    const RngKey key = key_of_slot(fp, slot);
    const uint32_t x = key.pixel % fp.width, y = key.pixel / fp.width;
    const uint32_t s_i = key.frame % fp.sqrt_spp, s_j = (key.frame / fp.sqrt_spp) % fp.sqrt_spp;
    const uint4 r = rng_draw(key, 0, kStreamCamera);
    const float offx = (static_cast<float>(s_i) + u01(r.x)) * fp.recip_sqrt_spp - 0.5f;
    const float offy = (static_cast<float>(s_j) + u01(r.y)) * fp.recip_sqrt_spp - 0.5f;
    const float fx = static_cast<float>(x) + offx, fy = static_cast<float>(y) + offy;
    F3 pc = {fp.cam.pixel00[0] + fx * fp.cam.pixel_delta_u[0] + fy * fp.cam.pixel_delta_v[0],
             fp.cam.pixel00[1] + fx * fp.cam.pixel_delta_u[1] + fy * fp.cam.pixel_delta_v[1],
             fp.cam.pixel00[2] + fx * fp.cam.pixel_delta_u[2] + fy * fp.cam.pixel_delta_v[2]};
    F3 c = {fp.cam.center[0], fp.cam.center[1], fp.cam.center[2]};
    if (fp.cam.defocus_angle > 0.0f) {
      // math::RandInUnitDisk (Math.hpp:34-41) sampled directly: uniform over the unit disk
      const uint4 l = rng_draw(key, 0, kStreamLens);
      float rr = sqrtf(u01(l.x)), s, co;
      sincospif(2.0f * u01(l.y), &s, &co);
      const float p0 = rr * co, p1 = rr * s;
      c = {c.x + p0 * fp.cam.defocus_disk_u[0] + p1 * fp.cam.defocus_disk_v[0],
           c.y + p0 * fp.cam.defocus_disk_u[1] + p1 * fp.cam.defocus_disk_v[1],
           c.z + p0 * fp.cam.defocus_disk_u[2] + p1 * fp.cam.defocus_disk_v[2]};
    }
    const F3 dir = normalize3(F3{pc.x - c.x, pc.y - c.y, pc.z - c.z});
    ray_o[slot] = make_float4(c.x, c.y, c.z, u01(r.z));  // .w = ray time
    ray_d[slot] = make_float4(dir.x, dir.y, dir.z, 0.0f);
    state[slot] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(slot));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) counters[0] = n_slots;
}

__device__ __forceinline__ int bin_of_material(const DeviceScene& S, int32_t material) {
  if (material < 0) return 0;
  const uint32_t type = __float_as_uint(__ldg(S.materials + 2 * material).x);
  switch (type) {
    case RT2_MAT_LAMBERTIAN: return 1;
    case RT2_MAT_TEXTURE: return 2;
    case RT2_MAT_METAL: return 3;
    case RT2_MAT_DIELECTRIC: return 4;
    case RT2_MAT_ISOTROPIC: return 5;
    default: return 0;  // diffuse light (terminal)
  }
}

// Warp-aggregated append of `value` to queue `bin` (one atomic per distinct bin per warp).
__device__ __forceinline__ void bin_push(uint32_t* __restrict__ counters, uint32_t* __restrict__ queues, uint32_t queue_stride,
                                         int bin, uint32_t value) {
  const unsigned active = __activemask();
  const unsigned peers = __match_any_sync(active, bin);
  const int leader = __ffs(peers) - 1;
  const unsigned lane = threadIdx.x & 31u;
  uint32_t base = 0;
  if (static_cast<int>(lane) == leader) base = atomicAdd(&counters[1 + bin], __popc(peers));
  base = __shfl_sync(peers, base, leader);
  const uint32_t pos = base + __popc(peers & ((1u << lane) - 1u));
  RT2_CHECK(pos < queue_stride && bin >= 0 && bin < kNumBins, kChkBin);
  queues[static_cast<size_t>(bin) * queue_stride + pos] = value;
}

// one buffer, kNumBins regions of `stride` entries
struct BinQueues {
  uint32_t* base;
  uint32_t stride;
  __host__ __device__ uint32_t* q(int bin) const { return base + static_cast<size_t>(bin) * stride; }
};

// Extend, part 1: closest SURFACE for every queued ray of this bounce (persistent warps, dynamic fetch; rt_trace.cuh).
//   kTravInline / kTravWorld: one item = one ray of the queue;  kTravInst: one item = one {ray, instance} entry.
// n_ptr: device-resident item count (queue size, or the bounce's entry counter); nullptr: n_fixed (rt2_intersect).
// cursor: the fetch cursor of the queue (zeroed with the other counters at batch start).
__device__ __forceinline__ void flush_trav_counters(const TravCounters& cnt, bool count_work, unsigned long long* __restrict__ totals) {
  if (count_work) {
    // warp-reduce, one atomic per warp and counter
    uint32_t v[4] = {cnt.box_pairs, cnt.spheres, cnt.quads, cnt.instances};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      uint32_t x = v[k];
      for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xFFFFFFFFu, x, off);
      if ((threadIdx.x & 31u) == 0u && x) atomicAdd(&totals[2 + k], static_cast<unsigned long long>(x));
    }
  }
  if (cnt.overflow) atomicAdd(&totals[6], 1ull);  // never expected: the builders bound the tree depth
}

#ifndef RT2_TRAV_MIN_BLOCKS
#define RT2_TRAV_MIN_BLOCKS 4  // 64 registers; 3 (85 registers, no spills, 24 warps per SM) measured slower — profiles/r02_notes.md
#endif
template <class M, bool kCount, int kMode, bool kQuant = false>
__global__ void __launch_bounds__(kBlock, RT2_TRAV_MIN_BLOCKS) k_traverse(const DeviceScene S, const uint32_t* __restrict__ n_ptr, uint32_t n_fixed,
                                                        uint32_t* __restrict__ cursor, const float4* __restrict__ ray_o,
                                                        const float4* __restrict__ ray_d, float tmin, float tmax,
                                                        const uint32_t* __restrict__ order, uint32_t sort_min_rays, uint4* trav,
                                                        const SplitIO io, unsigned long long* __restrict__ totals, int max_steps,
                                                        int fetch_threshold) {
  const uint32_t n = n_ptr ? *n_ptr : n_fixed;
  TravCounters cnt;
  __shared__ float s_ms[kMode == kTravUnified ? 14 * kBlock : 1];  // kTravUnified: world ray + model-space ray of the instance last met
  // queues below the threshold were not sorted (rt_sort.cuh): consume them in queue order
  const uint32_t* ord = (order != nullptr && n >= sort_min_rays) ? order : nullptr;
  // scene.hittable_list.Hit(scene, r, Interval{0.001, kInfinity}, rec)  (RayTracer.cpp:25)
  traverse_queue<M, kCount, kMode, kQuant>(S, n, ray_o, ray_d, tmin, tmax, cursor, ord, trav, io, cnt, max_steps, fetch_threshold, s_ms);
  flush_trav_counters(cnt, kCount, totals);
}

// Extend, part 1 over the 4-wide quantised tree (rt_wide.cuh): scenes whose binary tree does not fit the caches.
// n_ptr == nullptr: fixed ray count and caller-supplied interval (rt2_intersect).
template <class M, bool kCount>
__global__ void __launch_bounds__(kBlock, 4) k_traverse_wide(const DeviceScene S, const WideScene W, const uint32_t* __restrict__ n_ptr,
                                                            uint32_t n_fixed, uint32_t* __restrict__ cursor,
                                                            const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, float tmin,
                                                            float tmax, uint4* __restrict__ trav,
                                                            unsigned long long* __restrict__ totals, int max_steps) {
  const uint32_t n = n_ptr ? *n_ptr : n_fixed;
  TravCounters cnt;
  traverse_queue_wide<M, kCount, kFetchThreshold>(S, W, n, ray_o, ray_d, tmin, tmax, cursor, trav, cnt, max_steps);
  flush_trav_counters(cnt, kCount, totals);
}

// Extend, part 1 for tiny scenes (DeviceScene::flat_*): one thread per ray, uniform loops, no tree (traverse_flat).
template <class M>
__global__ void __launch_bounds__(kBlock) k_traverse_flat(const DeviceScene S, const uint32_t* __restrict__ counters, uint32_t n_fixed,
                                                          const float4* __restrict__ ray_o, const float4* __restrict__ ray_d, float tmin,
                                                          float tmax, uint4* __restrict__ trav) {
  const uint32_t n = counters ? counters[0] : n_fixed;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float4 o = ray_o[i], d = ray_d[i];
    const Closest best = traverse_flat<M>(S, make_f3(o), make_f3(d), o.w, tmin, tmax);
    trav[i] = make_uint4(__float_as_uint(best.t), best.prim, static_cast<uint32_t>(best.instance), 0u);
  }
}

// Constant media as a pass of their own (scenes whose media have multi-primitive boundaries, e.g. the box-bounded smoke of
// the Cornell-volume scene: 2 media x 6 exact quad tests per ray): uniform work with its own register budget and the
// axis-aligned quad path, instead of living inside the register-starved fused kernel.  A medium that scatters in front of
// the closest surface overwrites the ray's traversal record with {t, medium reference}; k_finish_shade<.., kMedia = false>
// turns either into the hit record.
template <class M>
__global__ void __launch_bounds__(kBlock) k_media(const DeviceScene S, const FrameParams fp, uint32_t bounce,
                                                  const uint32_t* __restrict__ counters, const float4* __restrict__ ray_o,
                                                  const float4* __restrict__ ray_d, const float4* __restrict__ state,
                                                  uint4* __restrict__ trav, const SplitIO io) {
  const uint32_t n = counters[0];
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float4 o = ray_o[i], d = ray_d[i];
    const uint4 tr = load_closest(S, trav, io, i);
    const RngKey key = key_of_slot(fp, __float_as_uint(state[i].w));
    Closest best{__uint_as_float(tr.x), tr.y, static_cast<int32_t>(tr.z)};
    int32_t mh = -1;
    media_sample<M, true>(S, make_f3(o), make_f3(d), o.w, 0.001f, key, bounce, best, mh);
    if (mh >= 0) trav[i] = make_uint4(__float_as_uint(best.t), (RT2_PRIM_MEDIUM << 28) | static_cast<uint32_t>(mh), 0xFFFFFFFFu, 0u);
  }
}

// Extend, part 2: constant media against the closest surface, the winner's hit record, and the push of the ray index
// into its material bin.  One thread per ray, uniform work.
template <class M>
__global__ void __launch_bounds__(kBlock) k_finish_hit(const DeviceScene S, const FrameParams fp, uint32_t bounce,
                                                       uint32_t* __restrict__ counters, const float4* __restrict__ ray_o,
                                                       const float4* __restrict__ ray_d, const float4* __restrict__ state,
                                                       const uint4* __restrict__ trav, const SplitIO io, float4* __restrict__ hit0,
                                                       float4* __restrict__ hit1, BinQueues bins) {
  const uint32_t n = counters[0];
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float4 o = ray_o[i], d = ray_d[i];
    const uint4 tr = load_closest(S, trav, io, i);
    const RngKey key = key_of_slot(fp, __float_as_uint(state[i].w));
    HitOut h;
    finish_hit<M>(S, make_f3(o), make_f3(d), o.w, 0.001f, Closest{__uint_as_float(tr.x), tr.y, static_cast<int32_t>(tr.z)}, key, bounce,
                  false, h);
    hit0[i] = make_float4(h.p.x, h.p.y, h.p.z, __uint_as_float(pack_uv16(h.u, h.v)));  // .w: (u, v); t is not needed downstream
    const uint32_t mbits = (h.material < 0) ? 0xFFFFFFFFu : (static_cast<uint32_t>(h.material) | (h.front_face ? 0x80000000u : 0u));
    hit1[i] = make_float4(h.n.x, h.n.y, h.n.z, __uint_as_float(mbits));
    bin_push(counters, bins.base, bins.stride, bin_of_material(S, h.material), i);
  }
}

// Fused extend part 2 + shade: constant media against the closest surface, the winner's hit record, and — for every
// material whose shading is cheap (everything except noise-textured ones) — Material::Scatter / Emit right here
// (RayTracer.cpp:25-44), with the scattered ray appended to the next bounce's queue by one warp-aggregated atomic.
// The hit record never goes to memory for these rays.  Materials flagged as deferred (rt2_material.pad != 0 on the
// device: their texture chain reaches a Perlin / marble texture, ~1 kflop per evaluation) get their record written to
// hit0 / hit1 and their index pushed into the material bin; k_shade_scatter / k_shade_terminal then run on those bins
// only, so one marble ray does not stall its 31 warp-mates.
//   counters[0] = queue size; counters[1..6] = deferred bins; next_counters[0] = rays emitted inline so far.
// kMedia: 0 = no media code (none in the scene, or k_media ran first), 1 = "simple" media (one-primitive world-space
// boundaries: book 2), 2 = general
// kImages: the scene has image textures (HitRecord::uv and the texel lookup are compiled in)
template <class M, int kMinBlocks = 4, int kMedia = 2, bool kImages = true, int kTile = kBlock>
__global__ void __launch_bounds__(kTile, kMinBlocks) k_finish_shade(const DeviceScene S, const FrameParams fp, uint32_t bounce, int emit_next,
                                                         uint32_t* __restrict__ counters, uint32_t* __restrict__ next_counters,
                                                         const float4* __restrict__ ray_o, const float4* __restrict__ ray_d,
                                                         const float4* __restrict__ state, const uint4* __restrict__ trav,
                                                         const SplitIO io, float4* __restrict__ hit0, float4* __restrict__ hit1, BinQueues bins,
                                                         float4* __restrict__ out_o, float4* __restrict__ out_d,
                                                         float4* __restrict__ out_state, uint32_t* __restrict__ sort_keys,
                                                         float4* __restrict__ radiance) {
  const uint32_t n = counters[0];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  constexpr int kWarps = kTile / 32;
  constexpr int kClasses = 5;  // lambertian, textured lambertian, metal, dielectric, isotropic
  __shared__ uint32_t s_off[2][kClasses][kWarps];
  __shared__ uint32_t s_base[2];
  int parity = 0;
  const uint32_t n_tiles = (n + kTile - 1) / kTile;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {  // block-uniform trip count
    const uint32_t i = tile * kTile + threadIdx.x;
    {
      // the block's next tile: pull its four input streams into L2 now (each warp stalls once per tile on these loads —
      // 12 % of the kernel's stall samples — and a DRAM miss costs twice an L2 hit; measured: long_scoreboard 3.45 -> 1.98
      // per issue, kernel -8 %)
      const uint32_t nx = i + gridDim.x * kTile;
      if (nx < n && (threadIdx.x & 7u) == 0u) {  // one request per 128-byte line of 8 float4
        prefetch_l2(ray_o + nx);
        prefetch_l2(ray_d + nx);
        prefetch_l2(trav + nx);
        prefetch_l2(state + nx);
      }
    }
    bool emit = false;
    int cls = 0;
    float4 no = make_float4(0, 0, 0, 0), nd = make_float4(0, 0, 0, 0), ns = make_float4(0, 0, 0, 0);
    if (i < n) {
      const float4 o = ray_o[i], d = ray_d[i];
      const uint4 tr = load_closest(S, trav, io, i);
      const float4 st = state[i];
      const uint32_t slot = __float_as_uint(st.w);
      RT2_CHECK(slot < bins.stride, kChkSlot);  // bins.stride = paths per batch = slots of the radiance buffer
      RT2_CHECK(o.x == o.x && d.x == d.x && st.x == st.x && __uint_as_float(tr.x) == __uint_as_float(tr.x), kChkNaN);
      const RngKey key = key_of_slot(fp, slot);
      HitOut h;
      if (kMedia) {
        finish_hit<M, kMedia == 1, kImages>(S, make_f3(o), make_f3(d), o.w, 0.001f, Closest{__uint_as_float(tr.x), tr.y, static_cast<int32_t>(tr.z)}, key,
                                   bounce, false, h);
      } else {
        // no media in the scene, or k_media ran before: a medium that won left its reference in the traversal record
        Closest best{__uint_as_float(tr.x), tr.y, static_cast<int32_t>(tr.z)};
        int32_t mh = -1;
        if (tr.y != RT2_PRIM_NONE && RT2_PRIM_TYPE(tr.y) == RT2_PRIM_MEDIUM) {
          mh = static_cast<int32_t>(RT2_PRIM_INDEX(tr.y));
          best.prim = RT2_PRIM_NONE;
        }
        finish_record<M, false, kImages>(S, make_f3(o), make_f3(d), o.w, best, mh, h);
      }
      if (h.material < 0) {
        // miss: T * background (RayTracer.cpp:25-27)
        radiance[slot] = make_float4(st.x * S.background[0], st.y * S.background[1], st.z * S.background[2], 0.0f);
      } else {
        RT2_CHECK(static_cast<uint32_t>(h.material) < S.n_materials, kChkMaterial);
        const float4 m0 = __ldg(S.materials + 2 * h.material), m1 = __ldg(S.materials + 2 * h.material + 1);
        const uint32_t type = __float_as_uint(m0.x);
        RT2_CHECK(type < RT2_MAT_INVALID, kChkMaterial);
        if (__float_as_uint(m1.w) != 0u) {
          // deferred (noise-textured): record + bin, shaded by the per-bin kernels
          hit0[i] = make_float4(h.p.x, h.p.y, h.p.z, __uint_as_float(pack_uv16(h.u, h.v)));
          hit1[i] = make_float4(h.n.x, h.n.y, h.n.z, __uint_as_float(static_cast<uint32_t>(h.material) | (h.front_face ? 0x80000000u : 0u)));
          bin_push(counters, bins.base, bins.stride, bin_of_material(S, h.material), i);
        } else if (type == RT2_MAT_DIFFUSE_LIGHT) {
          const F3 c = texture_value_simple<kImages>(S, __float_as_uint(m0.y), h.p, h.u, h.v);  // DiffuseLight::Emit (Material.cpp:71-74)
          radiance[slot] = make_float4(st.x * c.x, st.y * c.y, st.z * c.z, 0.0f);
        } else if (emit_next) {
          const uint4 r = rng_draw(key, bounce, kStreamScatter);
          F3 dir, att;
          if (type == RT2_MAT_DIELECTRIC) {
            att = scatter<RT2_MAT_DIELECTRIC>(S, m0, m1, make_f3(d), h.p, h.n, h.front_face, r, dir);
            cls = 3;
          } else {
            // every other material scatters around one uniform direction (math::RandUnitVec3): ONE copy of the sincos code
            const F3 u = unit_vector(u01(r.x), u01(r.y));
            if (type == RT2_MAT_METAL) {
              // Material.cpp:10-17: always scatters (no dot(scattered, normal) > 0 test)
              const F3 refl = normalize3(reflect3(make_f3(d), h.n));
              dir = {refl.x + m0.z * u.x, refl.y + m0.z * u.y, refl.z + m0.z * u.z};
              att = {m1.x, m1.y, m1.z};
              cls = 2;
            } else if (type == RT2_MAT_ISOTROPIC) {
              // Material.cpp:76-83
              dir = u;
              att = texture_value_simple<kImages>(S, __float_as_uint(m0.y), h.p, h.u, h.v);
              cls = 4;
            } else {
              // lambertian / textured lambertian, Material.cpp:47-69
              dir = {h.n.x + u.x, h.n.y + u.y, h.n.z + u.z};
              if (near_zero(dir)) dir = h.n;
              att = (type == RT2_MAT_LAMBERTIAN) ? F3{m1.x, m1.y, m1.z} : texture_value_simple<kImages>(S, __float_as_uint(m0.y), h.p, h.u, h.v);
              cls = (type == RT2_MAT_LAMBERTIAN) ? 0 : 1;
            }
          }
          emit = true;
          no = make_float4(h.p.x, h.p.y, h.p.z, o.w);
          nd = make_float4(dir.x, dir.y, dir.z, 0.0f);
          ns = make_float4(st.x * att.x, st.y * att.y, st.z * att.z, st.w);
        }
      }
    }
    // Append the tile's scattered rays to the next queue with ONE atomic per tile, grouped by material class inside the
    // tile's slice: runs of the next queue come from neighbouring pixels / queue positions and share a material, so the
    // warps of the next bounce start from similar places with similar direction distributions (a global stable
    // compaction by chained scan and plain per-warp atomics were both measured slower, profiles/r01_notes.md; staging the
    // tile in shared memory and flushing it one tile later, with the atomic's round trip hidden under the next tile's shading,
    // measured exactly the same: the two barriers cost what the slowest warp of the tile costs, profiles/r02_notes.md).
    uint32_t(*off)[kWarps] = s_off[parity];
    unsigned my_mask = 0;
#pragma unroll
    for (int c = 0; c < kClasses; c++) {
      const unsigned mc = __ballot_sync(0xFFFFFFFFu, emit && cls == c);
      if (lane == 0) off[c][warp] = __popc(mc);
      if (cls == c) my_mask = mc;
    }
    __syncthreads();
    if (warp == 0) {
      // exclusive scan of the kClasses x kWarps counts (class-major) by one warp: two entries per lane
      uint32_t* flat = &off[0][0];
      const uint32_t v0 = 2 * lane < kClasses * kWarps ? flat[2 * lane] : 0u;
      const uint32_t v1 = 2 * lane + 1 < kClasses * kWarps ? flat[2 * lane + 1] : 0u;
      uint32_t incl = v0 + v1;
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= static_cast<unsigned>(o)) incl += y;
      }
      const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
      uint32_t base = 0;
      if (lane == 0 && total) base = atomicAdd(&next_counters[0], total);
      const uint32_t excl = incl - (v0 + v1);
      if (2 * lane < kClasses * kWarps) flat[2 * lane] = excl;
      if (2 * lane + 1 < kClasses * kWarps) flat[2 * lane + 1] = excl + v0;
      if (lane == 0) s_base[parity] = base;
    }
    __syncthreads();
    if (emit) {
      const uint32_t dst = s_base[parity] + off[cls][warp] + __popc(my_mask & ((1u << lane) - 1u));
      RT2_CHECK(dst < bins.stride, kChkQueue);
      out_o[dst] = no;
      out_d[dst] = nd;
      out_state[dst] = ns;
#ifdef RT2_WITH_RAY_SORT
      if (sort_keys) sort_keys[dst] = ray_sort_key(S, make_f3(no), make_f3(nd));
#endif
    }
    parity ^= 1;  // the next tile uses the other copy of s_off / s_base: no third barrier
  }
}

// Paths that end here: miss -> background, emitter -> Emit (both faces).  radiance[slot] = throughput * colour.
__global__ void __launch_bounds__(kBlock) k_shade_terminal(const DeviceScene S, uint32_t* __restrict__ counters,
                                                           uint32_t* __restrict__ next_counters, const uint32_t* __restrict__ queue,
                                                           const float4* __restrict__ state, const float4* __restrict__ hit0,
                                                           const float4* __restrict__ hit1, float4* __restrict__ radiance, int fused) {
  const uint32_t n = counters[1];
  if (blockIdx.x == 0 && threadIdx.x == 0 && next_counters != nullptr) {
    // every ray in a scattering bin produces exactly one ray for the next bounce; in fused mode (k_finish_shade) the
    // bins hold the deferred rays only and next_counters[0] already counts the rays emitted inline
    const uint32_t binned = counters[2] + counters[3] + counters[4] + counters[5] + counters[6];
    next_counters[0] = fused ? next_counters[0] + binned : binned;
  }
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const uint32_t i = queue[j];
    const float4 st = state[i];
    const float4 h1 = hit1[i];
    const uint32_t mbits = __float_as_uint(h1.w);
    F3 c;
    if (mbits == 0xFFFFFFFFu) {
      c = {S.background[0], S.background[1], S.background[2]};
    } else {
      const float4 h0 = hit0[i];
      const float4 m0 = __ldg(S.materials + 2 * (mbits & 0x7FFFFFFFu));
      const float2 uv = unpack_uv16(__float_as_uint(h0.w));
      c = texture_value(S, __float_as_uint(m0.y), make_f3(h0), uv.x, uv.y);  // DiffuseLight::Emit (Material.cpp:71-74)
    }
    radiance[__float_as_uint(st.w)] = make_float4(st.x * c.x, st.y * c.y, st.z * c.z, 0.0f);
  }
}

template <int kType, int kBin>
__global__ void __launch_bounds__(kBlock) k_shade_scatter(const DeviceScene S, const FrameParams fp, uint32_t bounce,
                                                          const uint32_t* __restrict__ counters, const uint32_t* __restrict__ queue,
                                                          const float4* __restrict__ ray_o, const float4* __restrict__ ray_d,
                                                          const float4* __restrict__ state, const float4* __restrict__ hit0,
                                                          const float4* __restrict__ hit1, float4* __restrict__ out_o,
                                                          float4* __restrict__ out_d, float4* __restrict__ out_state,
                                                          uint32_t* __restrict__ sort_keys, const uint32_t* __restrict__ inline_count) {
  const uint32_t n = counters[1 + kBin];
  // fused mode: the deferred rays go behind the rays k_finish_shade emitted inline (k_shade_terminal, launched after
  // the scatter kernels, folds the bin sizes into next_counters[0])
  uint32_t base = inline_count ? inline_count[0] : 0u;
#pragma unroll
  for (int b = 1; b < kBin; b++) base += counters[1 + b];
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
    const uint32_t i = queue[j];
    const float4 st = state[i];
    const float4 h0 = hit0[i], h1 = hit1[i];
    const float4 d = ray_d[i];
    const float time = ray_o[i].w;
    const uint32_t mbits = __float_as_uint(h1.w);
    const uint32_t mat = mbits & 0x7FFFFFFFu;
    const float4 m0 = __ldg(S.materials + 2 * mat), m1 = __ldg(S.materials + 2 * mat + 1);
    const RngKey key = key_of_slot(fp, __float_as_uint(st.w));
    const uint4 r = rng_draw(key, bounce, kStreamScatter);
    F3 dir;
    const float2 uv = unpack_uv16(__float_as_uint(h0.w));
    const F3 att = scatter<kType>(S, m0, m1, make_f3(d), make_f3(h0), make_f3(h1), (mbits >> 31) != 0u, r, dir, uv.x, uv.y);
    const uint32_t dst = base + j;
    out_o[dst] = make_float4(h0.x, h0.y, h0.z, time);
    out_d[dst] = make_float4(dir.x, dir.y, dir.z, 0.0f);
    out_state[dst] = make_float4(st.x * att.x, st.y * att.y, st.z * att.z, st.w);
#ifdef RT2_WITH_RAY_SORT
    if (sort_keys) sort_keys[dst] = ray_sort_key(S, make_f3(h0), dir);
#endif
  }
}

// accumulation_data_[idx] += ray_color, one frame after the other (RayTracer.cpp:64) — fixed order, no atomics.
__global__ void __launch_bounds__(kBlock) k_accumulate(uint32_t pixels, uint32_t n_frames, const float4* __restrict__ radiance,
                                                       float4* __restrict__ accum, float4* __restrict__ accum_sq) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += stride) {
    float4 a = accum[p];
    float4 s = accum_sq ? accum_sq[p] : make_float4(0, 0, 0, 0);
    for (uint32_t f = 0; f < n_frames; f++) {
      const float4 v = radiance[static_cast<size_t>(f) * pixels + p];
      a.x += v.x;
      a.y += v.y;
      a.z += v.z;
      s.x += v.x * v.x;
      s.y += v.y * v.y;
      s.z += v.z * v.z;
    }
    accum[p] = a;
    if (accum_sq) accum_sq[p] = s;
  }
}

// NonConvertedPixels (RayTracer.cpp:105-112) and the RGBA8 preview (RayTracer.cpp:16-18,65-66)
__global__ void __launch_bounds__(kBlock) k_resolve(uint32_t pixels, float frames, const float4* __restrict__ accum,
                                                    float* __restrict__ mean_rgb, uchar4* __restrict__ rgba8) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < pixels; p += stride) {
    const float4 a = accum[p];
    const float mx = a.x / frames, my = a.y / frames, mz = a.z / frames;
    if (mean_rgb) {
      mean_rgb[3 * p + 0] = mx;
      mean_rgb[3 * p + 1] = my;
      mean_rgb[3 * p + 2] = mz;
    }
    if (rgba8) {
      // ToColor(glm::clamp(mean, 0, 1)): floor(c * 255.999) in double, then converted to uint8
      const double cx = floor(static_cast<double>(fminf(fmaxf(mx, 0.0f), 1.0f)) * 255.999);
      const double cy = floor(static_cast<double>(fminf(fmaxf(my, 0.0f), 1.0f)) * 255.999);
      const double cz = floor(static_cast<double>(fminf(fmaxf(mz, 0.0f), 1.0f)) * 255.999);
      rgba8[p] = make_uchar4(static_cast<unsigned char>(cx), static_cast<unsigned char>(cy), static_cast<unsigned char>(cz), 255);
    }
  }
}

// Multi-GPU read-out fused into one kernel: rank 0 reads the accumulators of its peers straight out of their
// HBM over NVLink (CUDA IPC mappings, P2P loads), adds them in rank order — so the N-GPU image has one defined summation
// order —, divides by the total frame count and writes the mean and the RGBA8 preview.  Replaces reduce + resolve.
constexpr int kMaxPeers = 16;
struct PeerAccums {
  const float4* p[kMaxPeers];
};
__global__ void __launch_bounds__(kBlock) k_resolve_peers(uint32_t pixels, float frames, PeerAccums peers, int n_ranks,
                                                          float* __restrict__ mean_rgb, uchar4* __restrict__ rgba8) {
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += stride) {
    float4 a = peers.p[0][i];
    for (int r = 1; r < n_ranks; r++) {
      const float4 v = peers.p[r][i];
      a.x += v.x;
      a.y += v.y;
      a.z += v.z;
    }
    const float mx = a.x / frames, my = a.y / frames, mz = a.z / frames;
    if (mean_rgb) {
      mean_rgb[3 * i + 0] = mx;
      mean_rgb[3 * i + 1] = my;
      mean_rgb[3 * i + 2] = mz;
    }
    if (rgba8) {
      const double cx = floor(static_cast<double>(fminf(fmaxf(mx, 0.0f), 1.0f)) * 255.999);
      const double cy = floor(static_cast<double>(fminf(fmaxf(my, 0.0f), 1.0f)) * 255.999);
      const double cz = floor(static_cast<double>(fminf(fmaxf(mz, 0.0f), 1.0f)) * 255.999);
      rgba8[i] = make_uchar4(static_cast<unsigned char>(cx), static_cast<unsigned char>(cy), static_cast<unsigned char>(cz), 255);
    }
  }
}

// rays traced in this batch = sum over bounces of the queue sizes
__global__ void k_batch_stats(const uint32_t* __restrict__ counters, uint32_t max_depth, unsigned long long* __restrict__ totals) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long rays = 0;
    for (uint32_t b = 0; b < max_depth; b++) rays += counters[b * kCounterStride];
    totals[0] += rays;
    totals[1] += counters[0];
  }
}

// Fixed-ray parity hook (rt2_intersect): the same traversal (k_traverse) followed by this record writer.
template <class M>
__global__ void __launch_bounds__(kBlock) k_finish_intersect(const DeviceScene S, const float4* __restrict__ ray_o,
                                                             const float4* __restrict__ ray_d, const uint4* __restrict__ trav,
                                                             const SplitIO io, uint32_t n, float tmin, int skip_media, uint32_t seed_lo,
                                                             uint32_t seed_hi, rt2_hit* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 o = ray_o[i], d = ray_d[i];
  const uint4 tr = load_closest(S, trav, io, i);
  RngKey key{i, 0u, seed_lo, seed_hi};
  HitOut h;
  finish_hit<M>(S, make_f3(o), make_f3(d), o.w, tmin, Closest{__uint_as_float(tr.x), tr.y, static_cast<int32_t>(tr.z)}, key, 0,
                skip_media != 0, h);
  rt2_hit r;
  r.point[0] = h.p.x, r.point[1] = h.p.y, r.point[2] = h.p.z;
  r.t = h.t;
  r.normal[0] = h.n.x, r.normal[1] = h.n.y, r.normal[2] = h.n.z;
  r.material = h.material;
  r.prim = (h.material < 0) ? RT2_PRIM_NONE : h.prim;
  r.instance = (h.material < 0) ? -1 : h.instance;
  r.front_face = h.front_face ? 1u : 0u;
  r.uv16 = (h.material < 0) ? 0u : pack_uv16(h.u, h.v);
  out[i] = r;
}

}  // namespace rt2dev

// ===================================================================================================================
// Host-side launcher
// ===================================================================================================================
namespace rt2 {

using namespace rt2dev;

#define RT2_CUDA(call)                                                                                   \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess) {                                                                            \
      err_ = std::string(#call) + " failed: " + cudaGetErrorString(e__);                                \
      return RT2_ERR_CUDA;                                                                               \
    }                                                                                                    \
  } while (0)

struct Renderer::Impl {
  DeviceScene ds{};
  // scene buffers
  void* d_spheres{nullptr};
  void* d_quads{nullptr};
  void* d_xforms{nullptr};
  void* d_instances{nullptr};
  void* d_media{nullptr};
  void* d_media_bounds{nullptr};
  size_t cap_media_bounds{0};
  void* d_images{nullptr};
  void* d_image_texels{nullptr};
  size_t cap_images{0}, cap_image_texels{0};
  void* d_flat_refs{nullptr};
  void* d_flat_offsets{nullptr};
  void* d_inst_bounds{nullptr};
  size_t cap_flat_refs{0}, cap_flat_offsets{0}, cap_inst_bounds{0};
  bool flat_mode{false};  // tiny scene: k_traverse_flat instead of the BVH walk
  bool quant_nodes{false};  // the walk reads the 32-byte quantised node pairs (rt_qnodes.cu) instead of the 64-byte float pairs
  void* d_qnodes{nullptr};
  size_t cap_qnodes{0};
  bool unified_mode{false};  // instances flattened into ONE world-space tree (kTravUnified): the default when the scene carries it
  void* d_inst_leaves{nullptr};
  size_t cap_inst_leaves{0};
  bool split_mode{false};  // instance split: k_traverse<kTravWorld> + k_traverse<kTravInst> (1..kMaxHoistedInstances instances)
  SplitIO split{};         // entry queue / per-entry winner / per-ray merge slot (allocated in Resize when split_mode)
  size_t split_capacity{0};  // entries the queue holds (= paths per batch x instances)
  bool split_media{false};  // media with multi-primitive boundaries: k_media pass + media-free fused kernel
  void* d_nodes4{nullptr};  // RT2_FLAG_WIDE_BVH: 4-wide quantised nodes collapsed from the device LBVH (rt_wide.cuh)
  size_t cap_nodes4{0};
  bool wide_mode{false};
  WideScene wide{};
  // pinned host staging: scene uploads are packed into it (true async H2D), read-backs land in it (true async D2H)
  char* h_stage{nullptr};
  size_t h_stage_cap{0}, h_stage_used{0};
  void* d_materials{nullptr};
  void* d_textures{nullptr};
  void* d_perlin{nullptr};
  void* d_prim_refs{nullptr};
  void* d_nodes{nullptr};
  size_t cap_spheres{0}, cap_quads{0}, cap_xforms{0}, cap_instances{0}, cap_media{0}, cap_materials{0}, cap_textures{0},
      cap_perlin{0}, cap_prim_refs{0}, cap_nodes{0};
  // wavefront state
  float4* ray_o[2]{nullptr, nullptr};
  float4* ray_d[2]{nullptr, nullptr};
  float4* state[2]{nullptr, nullptr};
  float4* hit0{nullptr};
  float4* hit1{nullptr};
  uint4* trav{nullptr};
  BinQueues bins{};
  uint32_t* counters{nullptr};
  float4* radiance{nullptr};
  float4* accum{nullptr};
  float4* accum_sq{nullptr};
  float* mean_rgb{nullptr};
  uchar4* rgba8{nullptr};
  unsigned long long* totals{nullptr};  // [0] rays [1] paths [2..5] work counters (box pairs, spheres, quads, instances)
  cudaStream_t stream{nullptr};
  cudaEvent_t ev_start{nullptr}, ev_stop{nullptr}, ev_done{nullptr};
  std::vector<cudaEvent_t> timing_events;  // start/stop pairs of rt2_update calls not yet folded into gpu_ms_total
  size_t timing_used{0};
  int grid_extend{0}, grid_stream{0};
  bool bin_present[kNumBins]{};
  std::vector<cudaEvent_t> prof_events;
  std::vector<int> prof_kind;
  // ray sort (rt_sort.cuh)
  uint32_t* sort_keys[2]{nullptr, nullptr};
  uint32_t* sort_vals[2]{nullptr, nullptr};
  uint32_t* sort_hist{nullptr};
  uint32_t* sort_bin_base{nullptr};
  int grid_sort{0};
  bool simple_media{false};  // every medium: one-primitive boundary, no instance chain (k_finish_shade<.., kMedia = 1>)
  int trav_max_steps{8};  // node steps per round of the while-while traversal; set per scene in UploadScene
  int trav_fetch_threshold{kFetchThreshold};
  bool fused{true};               // k_finish_shade instead of k_finish_hit + per-bin shade kernels
  bool deferred_present[kNumBins]{};  // fused mode: bins that can receive noise-textured (deferred) materials
  bool sort_enabled{false};
  uint32_t sort_min_rays{1u << 17};
  uint32_t sort_max_bounce{16};
  LbvhScratch lbvh_scratch;
  void* d_build_prims{nullptr};
  size_t cap_build_prims{0};
};

Renderer::Renderer() : impl_(new Impl) {}

Renderer::~Renderer() {
  if (!impl_) return;
  cudaSetDevice(cfg_.device);
  FreeState();
  Impl& m = *impl_;
  void* bufs[] = {m.d_spheres, m.d_quads, m.d_xforms, m.d_instances, m.d_media, m.d_materials, m.d_textures, m.d_perlin, m.d_prim_refs, m.d_nodes, m.d_media_bounds, m.d_flat_refs, m.d_flat_offsets, m.d_inst_bounds, m.d_images, m.d_image_texels, m.d_nodes4, m.d_inst_leaves, m.d_qnodes};
  for (void* b : bufs)
    if (b) cudaFree(b);
  if (m.totals) cudaFree(m.totals);
  if (m.h_stage) cudaFreeHost(m.h_stage);
  if (m.d_build_prims) cudaFree(m.d_build_prims);
  FreeLbvhScratch(&m.lbvh_scratch);
  if (m.ev_start) cudaEventDestroy(m.ev_start);
  if (m.ev_stop) cudaEventDestroy(m.ev_stop);
  if (m.ev_done) cudaEventDestroy(m.ev_done);
  for (cudaEvent_t e : m.prof_events) cudaEventDestroy(e);
  for (cudaEvent_t e : m.timing_events) cudaEventDestroy(e);
  if (m.stream) cudaStreamDestroy(m.stream);
  delete impl_;
}

void Renderer::FreeState() {
  Impl& m = *impl_;
  void* bufs[] = {m.ray_o[0], m.ray_o[1], m.ray_d[0], m.ray_d[1], m.state[0], m.state[1], m.hit0, m.hit1, m.trav, m.counters,
                  m.radiance, m.accum, m.accum_sq, m.mean_rgb, m.rgba8, m.sort_keys[0], m.sort_keys[1], m.sort_vals[0], m.sort_vals[1],
                  m.sort_hist, m.sort_bin_base};
  m.sort_keys[0] = m.sort_keys[1] = m.sort_vals[0] = m.sort_vals[1] = m.sort_hist = m.sort_bin_base = nullptr;
  for (void* b : bufs)
    if (b) cudaFree(b);
  if (m.bins.base) cudaFree(m.bins.base);
  m.bins.base = nullptr;
  if (m.split.entries) cudaFree(m.split.entries);
  if (m.split.entry_prim) cudaFree(m.split.entry_prim);
  if (m.split.inst_best) cudaFree(m.split.inst_best);
  m.split = SplitIO{};
  m.split_capacity = 0;
  state_ok_ = false;
  m.ray_o[0] = m.ray_o[1] = m.ray_d[0] = m.ray_d[1] = m.state[0] = m.state[1] = m.hit0 = m.hit1 = nullptr;
  m.trav = nullptr;
  m.counters = nullptr;
  m.radiance = m.accum = m.accum_sq = nullptr;
  m.mean_rgb = nullptr;
  m.rgba8 = nullptr;
}

int Renderer::Init(const HostScene& scene, const rt2_config& cfg) {
  cfg_ = cfg;
  Impl& m = *impl_;
  int n_dev = 0;
  RT2_CUDA(cudaGetDeviceCount(&n_dev));
  if (cfg.device < 0 || cfg.device >= n_dev) {
    err_ = "CUDA device " + std::to_string(cfg.device) + " not available (" + std::to_string(n_dev) + " devices)";
    return RT2_ERR_CUDA;
  }
  RT2_CUDA(cudaSetDevice(cfg.device));
  cudaDeviceProp prop;
  RT2_CUDA(cudaGetDeviceProperties(&prop, cfg.device));
  sm_count_ = prop.multiProcessorCount;
  RT2_CUDA(cudaStreamCreateWithFlags(&m.stream, cudaStreamNonBlocking));
  RT2_CUDA(cudaEventCreate(&m.ev_start));
  RT2_CUDA(cudaEventCreate(&m.ev_stop));
  RT2_CUDA(cudaMalloc(&m.totals, 8 * sizeof(unsigned long long)));
  RT2_CUDA(cudaMemset(m.totals, 0, 8 * sizeof(unsigned long long)));
  if (cfg_.max_depth < 1) cfg_.max_depth = 1;
  if (cfg_.samples_per_pixel < 1) cfg_.samples_per_pixel = 1;
  if (cfg_.frame_stride < 1) cfg_.frame_stride = 1;
  // persistent grids: resident blocks per SM x SM count
  int occ = 0;
  if (cfg_.flags & RT2_FLAG_FAST_MATH) {
    RT2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_traverse<FastMath, false, kTravInline>, kBlock, 0));
  } else {
    RT2_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_traverse<ExactMath, false, kTravInline>, kBlock, 0));
  }
  if (occ < 1) occ = 1;
  m.grid_extend = sm_count_ * occ;
  m.grid_stream = sm_count_ * 8;
  m.grid_sort = sm_count_ * 4;
  m.sort_enabled = (cfg_.flags & RT2_FLAG_SORT_RAYS) != 0;
#ifndef RT2_WITH_RAY_SORT
  if (m.sort_enabled) {
    err_ = "RT2_FLAG_SORT_RAYS: the ray sort measured as a net loss and is only compiled into EXPERIMENTS builds";
    return RT2_ERR_UNSUPPORTED;
  }
#endif
  m.sort_min_rays = static_cast<uint32_t>(TuneInt("RT2_SORT_MIN", m.sort_min_rays));
  m.sort_max_bounce = static_cast<uint32_t>(TuneInt("RT2_SORT_MAXB", m.sort_max_bounce));
  m.fused = (cfg_.flags & RT2_FLAG_NO_FUSED_SHADE) == 0;
  int rc = UploadScene(scene);
  if (rc != RT2_OK) return rc;
  int w = cfg.width > 0 ? cfg.width : scene.width;
  int h = cfg.height > 0 ? cfg.height : scene.height;
  return Resize(w, h);
}

constexpr size_t kStageMax = 512ull << 20;  // scenes beyond this are copied straight from pageable memory

static int EnsureStage(Renderer::Impl& m, size_t bytes, std::string* err);

// Copies `v` to the device buffer *dptr (grown on demand).  Small scenes go through the pinned staging arena so that the
// copy is a real asynchronous DMA from page-locked memory; the arena is recycled at the start of every UploadScene and
// the stream is synchronised at its end.
template <class T>
static int UploadBuf(Renderer::Impl& m, void** dptr, size_t* cap, const std::vector<T>& v, std::string* err) {
  size_t bytes = v.size() * sizeof(T);
  if (bytes > *cap || *dptr == nullptr) {
    if (*dptr) cudaFree(*dptr);
    size_t alloc = bytes > 0 ? bytes : 16;
    cudaError_t e = cudaMalloc(dptr, alloc);
    if (e != cudaSuccess) {
      *err = std::string("cudaMalloc(scene buffer) failed: ") + cudaGetErrorString(e);
      return RT2_ERR_CUDA;
    }
    *cap = alloc;
  }
  if (bytes) {
    const void* src = v.data();
    const size_t off = (m.h_stage_used + 255) & ~static_cast<size_t>(255);
    if (m.h_stage && off + bytes <= m.h_stage_cap) {
      std::memcpy(m.h_stage + off, v.data(), bytes);
      src = m.h_stage + off;
      m.h_stage_used = off + bytes;
    }
    cudaError_t e = cudaMemcpyAsync(*dptr, src, bytes, cudaMemcpyHostToDevice, m.stream);
    if (e != cudaSuccess) {
      *err = std::string("cudaMemcpyAsync(scene buffer) failed: ") + cudaGetErrorString(e);
      return RT2_ERR_CUDA;
    }
  }
  return RT2_OK;
}

static int EnsureStage(Renderer::Impl& m, size_t bytes, std::string* err) {
  if (bytes > kStageMax) return RT2_OK;  // too large to mirror in page-locked memory: UploadBuf falls back to pageable copies
  if (bytes > m.h_stage_cap) {
    if (m.h_stage) cudaFreeHost(m.h_stage);
    m.h_stage = nullptr;
    m.h_stage_cap = 0;
    void* p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes);
    if (e != cudaSuccess) {
      *err = std::string("cudaMallocHost(staging) failed: ") + cudaGetErrorString(e);
      return RT2_ERR_CUDA;
    }
    m.h_stage = static_cast<char*>(p);
    m.h_stage_cap = bytes;
  }
  return RT2_OK;
}

int Renderer::UploadScene(const HostScene& scene) {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  int rc;
#define UP(field, vec, cap)                                                   \
  rc = UploadBuf(m, &m.field, &m.cap, scene.vec, &err_);                      \
  if (rc != RT2_OK) return rc;
  {
    // staging arena: every buffer uploaded below (+ 256-byte alignment slack per buffer)
    size_t total = scene.spheres.size() * sizeof(rt2_sphere) + scene.quads.size() * sizeof(rt2_quad) + scene.xforms.size() * sizeof(rt2_xform) +
                   scene.instances.size() * sizeof(rt2_instance) + scene.media.size() * sizeof(rt2_medium) +
                   scene.materials.size() * sizeof(rt2_material) + scene.textures.size() * sizeof(rt2_texture) +
                   scene.perlin.size() * sizeof(rt2_perlin) + scene.prim_refs.size() * sizeof(uint32_t) +
                   scene.nodes.size() * sizeof(rt2_bvh_node) + scene.media_bounds.size() * sizeof(float) + scene.images.size() * sizeof(rt2_image) +
                   scene.image_texels.size() * sizeof(float) +
                   (kFlatMaxPrims + scene.instances.size() * 9 + 8) * sizeof(uint32_t) + 32 * 256;
    RT2_CUDA(cudaStreamSynchronize(m.stream));  // earlier copies out of the arena
    rc = EnsureStage(m, total, &err_);
    if (rc != RT2_OK) return rc;
    m.h_stage_used = 0;
  }
  const bool gpu_bvh = (cfg_.flags & RT2_FLAG_GPU_LBVH) != 0;
  if (!gpu_bvh && !scene.has_host_bvh) {
    err_ = "this scene was created without a host BVH: create the renderer with RT2_FLAG_GPU_LBVH";
    return RT2_ERR_STATE;
  }
  UP(d_spheres, spheres, cap_spheres)
  UP(d_quads, quads, cap_quads)
  UP(d_xforms, xforms, cap_xforms)
  {
    // device copy of the materials: pad != 0 marks materials whose texture chain reaches a noise texture (deferred
    // shading in k_finish_shade); the ABI struct the caller sees is unchanged
    std::vector<rt2_material> dev_mats = scene.materials;
    for (int b = 0; b < kNumBins; b++) m.deferred_present[b] = false;
    auto reaches_noise = [&](uint32_t tex) {
      std::vector<uint32_t> todo{tex};
      for (int guard = 0; guard < 64 && !todo.empty(); guard++) {
        const uint32_t t = todo.back();
        todo.pop_back();
        if (t >= scene.textures.size()) continue;
        const rt2_texture& tx = scene.textures[t];
        if (tx.type == RT2_TEX_NOISE) return true;
        if (tx.type == RT2_TEX_CHECKER) {
          todo.push_back(tx.even_tex_idx);
          todo.push_back(tx.odd_tex_idx);
        }
      }
      return false;
    };
    for (rt2_material& mat : dev_mats) {
      uint32_t flag = 0;
      const bool textured = mat.type == RT2_MAT_TEXTURE || mat.type == RT2_MAT_DIFFUSE_LIGHT || mat.type == RT2_MAT_ISOTROPIC;
      if (textured && reaches_noise(mat.tex_idx)) {
        flag = 1;
        m.deferred_present[mat.type == RT2_MAT_TEXTURE ? 2 : (mat.type == RT2_MAT_ISOTROPIC ? 5 : 0)] = true;
      }
      std::memcpy(&mat.pad, &flag, sizeof(flag));
    }
    rc = UploadBuf(m, &m.d_materials, &m.cap_materials, dev_mats, &err_);
    RT2_CUDA(cudaStreamSynchronize(m.stream));  // dev_mats is a temporary (needed when the arena is bypassed)
    if (rc != RT2_OK) return rc;
  }
  UP(d_textures, textures, cap_textures)
  UP(d_perlin, perlin, cap_perlin)
  UP(d_media_bounds, media_bounds, cap_media_bounds)
  UP(d_images, images, cap_images)
  UP(d_image_texels, image_texels, cap_image_texels)
  if (!gpu_bvh) {
    UP(d_instances, instances, cap_instances)
    UP(d_media, media, cap_media)
    UP(d_prim_refs, prim_refs, cap_prim_refs)
    {
      // device node format: left_first holds the traversal entry (make_leaf_entry: a one-primitive leaf carries the
      // primitive reference itself), count keeps the ABI values as count << 27 | left_first for rt2_read_bvh
      std::vector<rt2_bvh_node> dev_nodes = scene.nodes;
      for (rt2_bvh_node& nd : dev_nodes) {
        if (nd.count > 0) {
          if (nd.count > 16 || nd.left_first >= (1u << 26) || nd.left_first + nd.count > scene.prim_refs.size()) {
            err_ = "BVH leaf does not fit the 4-bit count / 26-bit index entry";
            return RT2_ERR_UNSUPPORTED;
          }
          const uint32_t first = nd.left_first, count = nd.count;
          nd.left_first = make_leaf_entry(first, count, scene.prim_refs[first]);
          nd.count = (count << 27) | first;
        }
      }
      rc = UploadBuf(m, &m.d_nodes, &m.cap_nodes, dev_nodes, &err_);
      if (rc != RT2_OK) return rc;
      RT2_CUDA(cudaStreamSynchronize(m.stream));  // dev_nodes is a temporary
    }
    n_node_pairs_ = static_cast<uint32_t>(scene.nodes.size() / 2);
    n_prim_refs_ = static_cast<uint32_t>(scene.prim_refs.size());
    {
      // depth of every host tree in node pairs: [0] TLAS, [1 + i] BLAS of instance i, (last) surfaces-only world tree
      auto depth_of = [&](uint32_t root) {
        uint32_t best = 0;
        std::vector<std::pair<uint32_t, uint32_t>> todo{{root, 1u}};
        while (!todo.empty()) {
          const auto [pair, depth] = todo.back();
          todo.pop_back();
          best = std::max(best, depth);
          for (int side = 0; side < 2; side++) {
            const rt2_bvh_node& nd = scene.nodes[2ull * pair + side];
            if (nd.count == 0 && nd.bmin[0] <= nd.bmax[0]) todo.emplace_back(nd.left_first, depth + 1);
          }
        }
        return best;
      };
      tree_depths_.clear();
      tree_depths_.push_back(depth_of(scene.tlas_root));
      for (const rt2_instance& in : scene.instances) tree_depths_.push_back(depth_of(in.blas_root));
      world_depth_ = scene.has_world_tlas ? depth_of(scene.tlas_world_root) : 0u;
      unified_depth_ = scene.has_unified_tlas ? depth_of(scene.tlas_unified_root) : 0u;
    }
    world_tree_ok_ = scene.has_world_tlas;
    world_root_ = scene.tlas_world_root;
    unified_tree_ok_ = scene.has_unified_tlas;
    unified_root_ = scene.tlas_unified_root;
    bvh_build_ms_ = 0;
    m.wide_mode = false;
    if (cfg_.flags & RT2_FLAG_WIDE_BVH) {
      err_ = "RT2_FLAG_WIDE_BVH needs RT2_FLAG_GPU_LBVH (the wide tree is collapsed from the device-built tree)";
      return RT2_ERR_INVALID_ARG;
    }
  } else {
    rc = BuildTreesOnDevice(scene);
    if (rc != RT2_OK) return rc;
  }
#undef UP
  scene_bytes_ = scene.spheres.size() * sizeof(rt2_sphere) + scene.quads.size() * sizeof(rt2_quad) +
                 scene.xforms.size() * sizeof(rt2_xform) + scene.instances.size() * sizeof(rt2_instance) +
                 scene.media.size() * sizeof(rt2_medium) + scene.materials.size() * sizeof(rt2_material) +
                 scene.textures.size() * sizeof(rt2_texture) + scene.perlin.size() * sizeof(rt2_perlin) +
                 scene.prim_refs.size() * sizeof(uint32_t) + scene.nodes.size() * sizeof(rt2_bvh_node);
  DeviceScene& d = m.ds;
  d.spheres = static_cast<const float4*>(m.d_spheres);
  d.quads = static_cast<const float4*>(m.d_quads);
  d.xforms = static_cast<const float4*>(m.d_xforms);
  d.instances = static_cast<const uint4*>(m.d_instances);
  d.media = static_cast<const uint4*>(m.d_media);
  d.media_bounds = static_cast<const float4*>(m.d_media_bounds);
  d.images = static_cast<const uint4*>(m.d_images);
  d.image_texels = static_cast<const float4*>(m.d_image_texels);
  d.n_images = static_cast<uint32_t>(scene.images.size());
  d.n_spheres = static_cast<uint32_t>(scene.spheres.size());
  d.n_quads = static_cast<uint32_t>(scene.quads.size());
  d.n_materials = static_cast<uint32_t>(scene.materials.size());
  d.n_textures = static_cast<uint32_t>(scene.textures.size());
  d.n_node_pairs = n_node_pairs_;
  d.n_prim_refs = n_prim_refs_;
  d.n_inst_leaves = static_cast<uint32_t>(scene.inst_leaves.size() / 2);
  d.materials = static_cast<const float4*>(m.d_materials);
  d.textures = static_cast<const float4*>(m.d_textures);
  d.perlin = static_cast<const rt2_perlin*>(m.d_perlin);
  d.prim_refs = static_cast<const uint32_t*>(m.d_prim_refs);
  d.nodes = static_cast<const float4*>(m.d_nodes);
  d.tlas_root = gpu_bvh ? 0u : scene.tlas_root;  // device build: tree 0 (the TLAS) starts at pair 0
  d.n_media = static_cast<uint32_t>(scene.media.size());
  d.n_instances = static_cast<uint32_t>(scene.instances.size());
  d.min_inv_scale = scene.min_inv_scale;
  for (int k = 0; k < 3; k++) d.background[k] = scene.background[k];
  // ray-sort grid: robust world bounds of the top-level leaves (a far-away giant such as a r = 1000 ground sphere must
  // not squeeze everything else into one cell): 0.5 / 99.5 percentiles of the box corners, 32 cells per axis
  for (int k = 0; k < 3; k++) {
    d.sort_lo[k] = 0.0f;
    d.sort_scale[k] = 0.0f;
  }
  if (!scene.tree_prims.empty() && !scene.tree_prims[0].empty()) {
    const std::vector<BuildPrim>& tp = scene.tree_prims[0];
    const size_t cut = tp.size() / 200;
    std::vector<float> lo(tp.size()), hi(tp.size());
    for (int k = 0; k < 3; k++) {
      for (size_t i = 0; i < tp.size(); i++) lo[i] = tp[i].bmin[k], hi[i] = tp[i].bmax[k];
      std::nth_element(lo.begin(), lo.begin() + cut, lo.end());
      std::nth_element(hi.begin(), hi.end() - 1 - cut, hi.end());
      const float a = lo[cut], b = hi[hi.size() - 1 - cut];
      d.sort_lo[k] = a;
      d.sort_scale[k] = b > a ? 32.0f / (b - a) : 0.0f;
    }
  }
  // which material bins can ever be non-empty
  for (int b = 0; b < kNumBins; b++) m.bin_present[b] = false;
  m.bin_present[0] = true;
  for (const rt2_material& mat : scene.materials) {
    switch (mat.type) {
      case RT2_MAT_LAMBERTIAN: m.bin_present[1] = true; break;
      case RT2_MAT_TEXTURE: m.bin_present[2] = true; break;
      case RT2_MAT_METAL: m.bin_present[3] = true; break;
      case RT2_MAT_DIELECTRIC: m.bin_present[4] = true; break;
      case RT2_MAT_ISOTROPIC: m.bin_present[5] = true; break;
      default: break;
    }
  }
  m.split_media = false;
  m.simple_media = !scene.media.empty();
  for (const rt2_medium& md : scene.media) {
    m.split_media = m.split_media || md.boundary_count > 1;
    m.simple_media = m.simple_media && md.boundary_count == 1 && md.chain_len == 0;
  }
  if (TuneInt("RT2_SIMPLE_MEDIA", 1) == 0) m.simple_media = false;
  if (TuneInt("RT2_SPLIT_MEDIA", -1) >= 0) m.split_media = !scene.media.empty() && TuneInt("RT2_SPLIT_MEDIA", 0) != 0;
  // conservative world boxes of the instances (flat mode and instance split)
  d.inst_bounds = nullptr;
  if (!scene.inst_bounds.empty()) {
    if (scene.inst_bounds.size() != scene.instances.size() * 8) {
      err_ = "scene carries inconsistent instance bounds";
      return RT2_ERR_STATE;
    }
    rc = UploadBuf(m, &m.d_inst_bounds, &m.cap_inst_bounds, scene.inst_bounds, &err_);
    if (rc != RT2_OK) return rc;
    d.inst_bounds = static_cast<const float4*>(m.d_inst_bounds);
  }
  // flat mode: list every leaf primitive by space when the scene is tiny
  m.flat_mode = false;
  d.flat_refs = nullptr;
  d.flat_offsets = nullptr;
  uint32_t flat_max = static_cast<uint32_t>(TuneInt("RT2_FLAT_MAX", kFlatMaxPrims));
  if (TuneInt("RT2_FLAT", 1) == 0 || (cfg_.flags & RT2_FLAG_NO_FLAT_EXTEND)) flat_max = 0u;
  // (an explicit RT2_FLAG_GPU_LBVH asks for the device-built trees: keep the BVH walk)
  if (!gpu_bvh && flat_max > 0 && scene.tree_prims.size() == scene.instances.size() + 1 &&
      (scene.instances.empty() || d.inst_bounds != nullptr)) {
    std::vector<uint32_t> refs, offsets;
    bool ok = true;
    offsets.push_back(0);
    for (const BuildPrim& bp : scene.tree_prims[0]) {
      if (RT2_PRIM_TYPE(bp.ref) == RT2_PRIM_INSTANCE) {
        if (RT2_PRIM_INDEX(bp.ref) >= scene.instances.size()) {
          ok = false;
          break;
        }
      } else {
        refs.push_back(bp.ref);
      }
    }
    offsets.push_back(static_cast<uint32_t>(refs.size()));
    for (size_t j = 0; ok && j < scene.instances.size(); j++) {
      for (const BuildPrim& bp : scene.tree_prims[1 + j]) {
        if (RT2_PRIM_TYPE(bp.ref) == RT2_PRIM_INSTANCE) ok = false;  // nested instance references are flattened by the host; never expected
        refs.push_back(bp.ref);
      }
      offsets.push_back(static_cast<uint32_t>(refs.size()));
    }
    if (ok && !refs.empty() && refs.size() <= flat_max) {
      rc = UploadBuf(m, &m.d_flat_refs, &m.cap_flat_refs, refs, &err_);
      if (rc != RT2_OK) return rc;
      rc = UploadBuf(m, &m.d_flat_offsets, &m.cap_flat_offsets, offsets, &err_);
      if (rc != RT2_OK) return rc;
      RT2_CUDA(cudaStreamSynchronize(m.stream));  // the host vectors are temporaries
      d.flat_refs = static_cast<const uint32_t*>(m.d_flat_refs);
      d.flat_offsets = static_cast<const uint32_t*>(m.d_flat_offsets);
      m.flat_mode = true;
    }
  }
  // instance split: the few instances are hoisted out of the world tree (rt_trace.cuh, kTravWorld / kTravInst)
  // How instances are walked: unified world tree (default when the scene carries one), two-pass split (RT2_FLAG_INSTANCE_SPLIT,
  // 1..4 instances) or inline TLAS -> BLAS (RT2_FLAG_INSTANCES_INLINE, and every scene too big to flatten).
  const bool want_inline = (cfg_.flags & RT2_FLAG_INSTANCES_INLINE) != 0;
  const bool want_split = (cfg_.flags & RT2_FLAG_INSTANCE_SPLIT) != 0 && !want_inline;
  const bool was_split = m.split_mode;
  m.split_mode = !m.flat_mode && !m.wide_mode && want_split && world_tree_ok_ && !scene.instances.empty() &&
                 scene.instances.size() <= kMaxHoistedInstances && d.inst_bounds != nullptr;
  m.unified_mode = !m.flat_mode && !m.wide_mode && !want_inline && !m.split_mode && unified_tree_ok_ && !scene.instances.empty();
  d.inst_leaves = nullptr;
  d.tlas_unified_root = d.tlas_root;
  if (m.unified_mode) {
    rc = UploadBuf(m, &m.d_inst_leaves, &m.cap_inst_leaves, scene.inst_leaves, &err_);
    if (rc != RT2_OK) return rc;
    d.inst_leaves = static_cast<const uint2*>(m.d_inst_leaves);
    d.tlas_unified_root = unified_root_;
  }
  // Round length (node steps before the warp's leaf phase) and refill threshold (busy lanes below which the warp fetches new rays)
  // of the while-while walk.  Measured optima (profiles/r02_notes.md): the unified walk, whose leaf phase serves world and
  // instanced primitives with one test, wants shorter rounds — 6 / 24: book 2 6 572 Mrays/s against 6 300 at 8 / 20; the plain walk
  // keeps 8 / 20 (book 1 10 032 against 9 978, 1 M spheres 2 901 against 2 780).
  m.trav_max_steps = static_cast<int>(TuneInt("RT2_TRAV_STEPS", m.unified_mode ? 6 : 8));
  m.trav_fetch_threshold = static_cast<int>(TuneInt("RT2_TRAV_FETCH", m.unified_mode ? 24 : kFetchThreshold));
  d.n_hoisted = m.split_mode ? static_cast<uint32_t>(scene.instances.size()) : 0u;
  d.tlas_world_root = m.split_mode ? world_root_ : d.tlas_root;
  {
    // The traversal keeps one stack entry per level of the tree it walks and does no bounds checks (rt_trace.cuh): refuse a
    // scene whose trees do not fit instead of dropping sub-trees or writing outside the stack.
    const size_t n_inst = scene.instances.size();
    uint32_t blas_max = 0;
    for (size_t i = 0; i < n_inst && 1 + i < tree_depths_.size(); i++) blas_max = std::max(blas_max, tree_depths_[1 + i]);
    const uint32_t tlas = tree_depths_.empty() ? 0u : tree_depths_[0];
    const uint32_t world = world_tree_ok_ ? world_depth_ : 0u;
    max_stack_need_ = m.unified_mode ? unified_depth_ : (m.split_mode ? std::max(world, blas_max) : (n_inst ? tlas + 1 + blas_max : tlas));
    if (max_stack_need_ > static_cast<uint32_t>(kStackSize - 1) && !m.flat_mode && !m.wide_mode) {
      err_ = "BVH too deep for the traversal stack: needs " + std::to_string(max_stack_need_) + " entries, the stack holds " +
             std::to_string(kStackSize - 1);
      return RT2_ERR_UNSUPPORTED;
    }
  }
  // 32-byte node pairs for the unified and the inline walk (rt_qnodes.cu): used when the measured growth of the boxes' surface
  // area — the expected number of extra node visits — stays small.  Scenes whose leaves are tiny against the scene's extent
  // (BASELINE config 5: 10 M spheres of radius 0.2 over kilometres) keep the float nodes.
  m.quant_nodes = false;
  d.qnodes = nullptr;
  node_inflation_ = 0.0f;
  if (!m.flat_mode && !m.wide_mode && !m.split_mode && !(cfg_.flags & RT2_FLAG_FLOAT_NODES) && n_node_pairs_ > 0 &&
      TuneInt("RT2_QUANT_NODES", 1) != 0) {
    std::vector<uint32_t> roots;
    if (m.unified_mode) {
      roots.push_back(d.tlas_unified_root);
    } else {
      roots.push_back(d.tlas_root);
      std::vector<rt2_instance> inst(scene.instances.size());
      if (!inst.empty()) {
        // (the device copy: a device build lays the trees out itself)
        RT2_CUDA(cudaStreamSynchronize(m.stream));
        RT2_CUDA(cudaMemcpy(inst.data(), m.d_instances, inst.size() * sizeof(rt2_instance), cudaMemcpyDeviceToHost));
      }
      for (const rt2_instance& in : inst) roots.push_back(in.blas_root);
    }
    if (roots.size() <= 4096) {
      const size_t bytes = static_cast<size_t>(n_node_pairs_) * 32;
      if (bytes > m.cap_qnodes || m.d_qnodes == nullptr) {
        if (m.d_qnodes) cudaFree(m.d_qnodes);
        m.d_qnodes = nullptr;
        m.cap_qnodes = 0;
        RT2_CUDA(cudaMalloc(&m.d_qnodes, bytes));
        m.cap_qnodes = bytes;
      }
      NodeGrid grid{};
      rc = QuantiseNodesOnDevice(m.d_nodes, n_node_pairs_, roots.data(), static_cast<uint32_t>(roots.size()), m.d_qnodes, &grid, m.stream,
                                 &launches_, &err_);
      if (rc != RT2_OK) return rc;
      node_inflation_ = static_cast<float>(grid.inflation);
      const float limit = TuneFloat("RT2_QUANT_MAX_INFLATION", kQuantMaxInflation);
      if (grid.usable && grid.inflation <= limit) {
        m.quant_nodes = true;
        d.qnodes = static_cast<const uint4*>(m.d_qnodes);
        for (int k = 0; k < 3; k++) d.q_base[k] = grid.base[k], d.q_ext[k] = grid.ext[k];
      }
    }
  }
  if (m.split_mode && (!was_split || !m.split.entries) && width_ > 0) {
    rc = AllocSplitState();  // a re-upload switched the renderer into split mode after Resize
    if (rc != RT2_OK) return rc;
  }
  cam_params_ = scene.cam;
  n_textures_ = static_cast<uint32_t>(scene.textures.size());
  RT2_CUDA(cudaStreamSynchronize(m.stream));  // the staging arena (and any temporary above) may be reused from here on
  return RT2_OK;
}

// RT2_FLAG_GPU_LBVH: every tree (world TLAS + one BLAS per instance) is built on the device from its leaf records.
// Pair / reference index spaces are laid out here: media boundary references first, then the trees in order.
int Renderer::BuildTreesOnDevice(const HostScene& scene) {
  Impl& m = *impl_;
  if (scene.tree_prims.size() != scene.instances.size() + 1) {
    err_ = "scene carries no BVH build input";
    return RT2_ERR_STATE;
  }
  // media boundary primitive lists keep their own slice of prim_refs
  std::vector<uint32_t> prefix;
  std::vector<rt2_medium> media = scene.media;
  for (rt2_medium& md : media) {
    const uint32_t first = static_cast<uint32_t>(prefix.size());
    for (uint32_t i = 0; i < md.boundary_count; i++) prefix.push_back(scene.prim_refs[md.boundary_first + i]);
    md.boundary_first = first;
  }
  std::vector<rt2_instance> instances = scene.instances;
  // trees: [0] world TLAS (instances as leaves), [1 + i] BLAS of instance i, and — for the instance split — one more
  // world tree over the surfaces only
  std::vector<const std::vector<BuildPrim>*> trees;
  for (const std::vector<BuildPrim>& tp : scene.tree_prims) trees.push_back(&tp);
  std::vector<BuildPrim> surfaces;
  const bool want_world_tree = !scene.instances.empty() && scene.instances.size() <= kMaxHoistedInstances;
  if (want_world_tree) {
    for (const BuildPrim& bp : scene.tree_prims[0])
      if (RT2_PRIM_TYPE(bp.ref) != RT2_PRIM_INSTANCE) surfaces.push_back(bp);
    trees.push_back(&surfaces);
  }
  const bool want_unified_tree = scene.has_unified_tlas && !scene.unified_prims.empty();
  if (want_unified_tree) trees.push_back(&scene.unified_prims);
  const size_t n_trees = trees.size();
  const size_t world_slot = want_world_tree ? scene.tree_prims.size() : 0, unified_slot = want_unified_tree ? n_trees - 1 : 0;
  std::vector<uint32_t> pair_base(n_trees), ref_base(n_trees);
  uint64_t pairs = 0, refs = prefix.size();
  size_t max_n = 0;
  for (size_t k = 0; k < n_trees; k++) {
    const size_t n = trees[k]->size();
    pair_base[k] = static_cast<uint32_t>(pairs);
    ref_base[k] = static_cast<uint32_t>(refs);
    pairs += n > 1 ? n - 1 : 1;
    refs += n;
    max_n = n > max_n ? n : max_n;
    if (k > 0 && k <= scene.instances.size()) instances[k - 1].blas_root = pair_base[k];
  }
  world_tree_ok_ = want_world_tree;
  world_root_ = want_world_tree ? pair_base[world_slot] : 0u;
  unified_tree_ok_ = want_unified_tree;
  unified_root_ = want_unified_tree ? pair_base[unified_slot] : 0u;
  if (pairs >= 0x7FFFFFF0ull || refs >= 0x03FFFFFFull) {
    err_ = "scene too large for the 26-bit primitive / 31-bit node index fields";
    return RT2_ERR_UNSUPPORTED;
  }
  int rc;
  rc = UploadBuf(m, &m.d_instances, &m.cap_instances, instances, &err_);
  if (rc != RT2_OK) return rc;
  rc = UploadBuf(m, &m.d_media, &m.cap_media, media, &err_);
  if (rc != RT2_OK) return rc;
  auto ensure = [&](void** ptr, size_t* cap, size_t bytes) -> int {
    if (bytes > *cap || *ptr == nullptr) {
      if (*ptr) cudaFree(*ptr);
      *ptr = nullptr;
      *cap = 0;
      RT2_CUDA(cudaMalloc(ptr, bytes > 0 ? bytes : 16));
      *cap = bytes > 0 ? bytes : 16;
    }
    return RT2_OK;
  };
  rc = ensure(&m.d_nodes, &m.cap_nodes, pairs * 64);
  if (rc != RT2_OK) return rc;
  rc = ensure(&m.d_prim_refs, &m.cap_prim_refs, refs * 4);
  if (rc != RT2_OK) return rc;
  rc = ensure(&m.d_build_prims, &m.cap_build_prims, max_n * sizeof(BuildPrim));
  if (rc != RT2_OK) return rc;
  if (!prefix.empty()) RT2_CUDA(cudaMemcpyAsync(m.d_prim_refs, prefix.data(), prefix.size() * 4, cudaMemcpyHostToDevice, m.stream));
  uint32_t* d_depths = nullptr;
  RT2_CUDA(cudaMalloc(&d_depths, n_trees * sizeof(uint32_t)));
  RT2_CUDA(cudaMemsetAsync(d_depths, 0, n_trees * sizeof(uint32_t), m.stream));
  RT2_CUDA(cudaEventRecord(m.ev_start, m.stream));
  // (Chaining "giant" primitives — the r = 1e5 ground sphere — at the top of the radix tree instead of filing them among their
  // Morton neighbours was tried and measured: 34.0 vs 30.0 node pairs per ray at 1 M spheres, 45.2 vs 48.0 at 10 M; not kept.)
  for (size_t k = 0; k < n_trees; k++) {
    const std::vector<BuildPrim>& tp = *trees[k];
    if (!tp.empty()) {
      RT2_CUDA(cudaMemcpyAsync(m.d_build_prims, tp.data(), tp.size() * sizeof(BuildPrim), cudaMemcpyHostToDevice, m.stream));
    }
    rc = BuildLbvhOnDevice(static_cast<const BuildPrim*>(m.d_build_prims), static_cast<uint32_t>(tp.size()), pair_base[k], ref_base[k], m.d_nodes,
                           static_cast<uint32_t*>(m.d_prim_refs), &m.lbvh_scratch, m.stream, &launches_, d_depths + k,
                           (cfg_.flags & RT2_FLAG_LBVH_PLOC) != 0, &err_);
    if (rc != RT2_OK) {
      cudaFree(d_depths);
      return rc;
    }
    if (k + 1 < n_trees) RT2_CUDA(cudaStreamSynchronize(m.stream));  // d_build_prims is reused by the next tree
  }
  m.wide_mode = false;
  if (cfg_.flags & RT2_FLAG_WIDE_BVH) {
    if (!scene.instances.empty()) {
      err_ = "RT2_FLAG_WIDE_BVH supports scenes without instances (the wide tree is a single world-space tree)";
      return RT2_ERR_UNSUPPORTED;
    }
    const uint32_t n_pairs = static_cast<uint32_t>(pairs);
    rc = ensure(&m.d_nodes4, &m.cap_nodes4, static_cast<size_t>(n_pairs) * 64);
    if (rc != RT2_OK) return rc;
    k_wide_collapse<<<(n_pairs + 255) / 256, 256, 0, m.stream>>>(static_cast<const float4*>(m.d_nodes), n_pairs, static_cast<uint4*>(m.d_nodes4));
    launches_++;
    m.wide.nodes4 = static_cast<const uint4*>(m.d_nodes4);
    m.wide.root = pair_base[0];
    m.wide_mode = scene.tree_prims[0].size() >= 2;  // a one-leaf tree has no node pair to collapse
  }
  RT2_CUDA(cudaEventRecord(m.ev_stop, m.stream));
  tree_depths_.assign(n_trees, 0u);
  {
    const cudaError_t e = cudaMemcpyAsync(tree_depths_.data(), d_depths, n_trees * sizeof(uint32_t), cudaMemcpyDeviceToHost, m.stream);
    const cudaError_t e2 = cudaStreamSynchronize(m.stream);
    cudaFree(d_depths);
    RT2_CUDA(e);
    RT2_CUDA(e2);
  }
  world_depth_ = want_world_tree ? tree_depths_[world_slot] : 0u;
  unified_depth_ = want_unified_tree ? tree_depths_[unified_slot] : 0u;
  float ms = 0;
  cudaEventElapsedTime(&ms, m.ev_start, m.ev_stop);
  bvh_build_ms_ = ms;
  n_node_pairs_ = static_cast<uint32_t>(pairs);
  n_prim_refs_ = static_cast<uint32_t>(refs);
  return RT2_OK;
}

int Renderer::ReadBvh(rt2_bvh_node* nodes, size_t max_nodes, uint32_t* prim_refs, size_t max_refs, uint32_t* n_pairs, uint32_t* n_refs,
                      uint32_t* tlas_root) {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  *n_pairs = n_node_pairs_;
  *n_refs = n_prim_refs_;
  *tlas_root = m.ds.tlas_root;
  RT2_CUDA(cudaStreamSynchronize(m.stream));
  if (nodes) {
    if (max_nodes < 2ull * n_node_pairs_) {
      err_ = "node buffer too small";
      return RT2_ERR_INVALID_ARG;
    }
    RT2_CUDA(cudaMemcpy(nodes, m.d_nodes, 2ull * n_node_pairs_ * sizeof(rt2_bvh_node), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < 2ull * n_node_pairs_; i++) {  // device format -> ABI format (see UploadScene)
      if (nodes[i].count > 0) {
        const uint32_t packed = nodes[i].count;
        nodes[i].left_first = packed & 0x07FFFFFFu;
        nodes[i].count = packed >> 27;
      }
    }
  }
  if (prim_refs) {
    if (max_refs < n_prim_refs_) {
      err_ = "prim_refs buffer too small";
      return RT2_ERR_INVALID_ARG;
    }
    if (n_prim_refs_) RT2_CUDA(cudaMemcpy(prim_refs, m.d_prim_refs, static_cast<size_t>(n_prim_refs_) * 4, cudaMemcpyDeviceToHost));
  }
  return RT2_OK;
}

// Entry queue of the instance split: every ray can produce one entry per hoisted instance.
int Renderer::AllocSplitState() {
  Impl& m = *impl_;
  if (m.split.entries) cudaFree(m.split.entries);
  if (m.split.entry_prim) cudaFree(m.split.entry_prim);
  if (m.split.inst_best) cudaFree(m.split.inst_best);
  m.split = SplitIO{};
  m.split_capacity = 0;
  const size_t N = static_cast<size_t>(width_) * height_ * static_cast<size_t>(frames_per_batch_);
  const size_t cap = N * m.ds.n_hoisted;
  if (cap == 0) return RT2_OK;
  if (cap >= 0xFFFFFFF0ull) {
    err_ = "batch too large for the instance-split entry queue";
    return RT2_ERR_INVALID_ARG;
  }
  RT2_CUDA(cudaMalloc(&m.split.entries, cap * sizeof(uint2)));
  RT2_CUDA(cudaMalloc(&m.split.entry_prim, cap * sizeof(uint32_t)));
  RT2_CUDA(cudaMalloc(&m.split.inst_best, N * sizeof(unsigned long long)));
  m.split_capacity = cap;
  m.split.capacity = static_cast<uint32_t>(cap);
  return RT2_OK;
}

int Renderer::Resize(int w, int h) {
  Impl& m = *impl_;
  if (w <= 0 || h <= 0) {
    err_ = "invalid dims";
    return RT2_ERR_INVALID_ARG;
  }
  RT2_CUDA(cudaSetDevice(cfg_.device));
  RT2_CUDA(cudaStreamSynchronize(m.stream));
  FreeState();
  pending_frames_ = 0;
  width_ = w;
  height_ = h;
  // camera for the new dims (Camera::SetDims + Update)
  HostScene tmp;
  tmp.cam = cam_params_;
  tmp.width = w;
  tmp.height = h;
  tmp.UpdateCamera();
  camera_ = tmp.camera_block;
  int rc = AllocState();
  if (rc != RT2_OK) {
    // a failed allocation (e.g. out of memory at a huge size) must not leave a half-built renderer behind: every later
    // call reports RT2_ERR_STATE until a Resize succeeds
    FreeState();
    cudaGetLastError();
    width_ = height_ = 0;
    return rc;
  }
  state_ok_ = true;
  return Reset();
}

int Renderer::AllocState() {
  Impl& m = *impl_;
  const int w = width_, h = height_;
  const size_t P = static_cast<size_t>(w) * h;
  int F = cfg_.frames_per_batch;
  if (F <= 0) {
    // ~192 M paths in flight (~35 GB of wavefront state out of 180 GB of HBM): besides amortising the launches of the 50
    // bounces, big batches keep the LATE queues big — on book 2 a third of all rays are traced at bounce 12 or later, where
    // a 24 M-path batch leaves ~1 M-ray queues that run at half the per-ray speed (measured: 4 914 / 5 145 / 5 270 Mrays/s at
    // 64 / 128 / 256 frames per batch of 600 x 600)
    // (r02: 512 frames per batch of 600 x 600 measured +2 % over 256 — 5 907 vs 5 783 Mrays/s — and costs 35 GB of the 180 GB)
    F = static_cast<int>((192u * 1024u * 1024u + P - 1) / P);
    if (F < 1) F = 1;
    if (F > 512) F = 512;
  }
  frames_per_batch_ = F;
  const size_t N = P * static_cast<size_t>(F);
  if (N >= 0xFFFFFFF0ull) {
    err_ = "batch too large";
    return RT2_ERR_INVALID_ARG;
  }
  for (int i = 0; i < 2; i++) {
    RT2_CUDA(cudaMalloc(&m.ray_o[i], N * sizeof(float4)));
    RT2_CUDA(cudaMalloc(&m.ray_d[i], N * sizeof(float4)));
    RT2_CUDA(cudaMalloc(&m.state[i], N * sizeof(float4)));
  }
  RT2_CUDA(cudaMalloc(&m.hit0, N * sizeof(float4)));
  RT2_CUDA(cudaMalloc(&m.hit1, N * sizeof(float4)));
  RT2_CUDA(cudaMalloc(&m.trav, N * sizeof(uint4)));
  RT2_CUDA(cudaMalloc(&m.bins.base, N * kNumBins * sizeof(uint32_t)));
  m.bins.stride = static_cast<uint32_t>(N);
  RT2_CUDA(cudaMalloc(&m.counters, static_cast<size_t>(cfg_.max_depth + 1) * kCounterStride * sizeof(uint32_t)));
  RT2_CUDA(cudaMalloc(&m.radiance, N * sizeof(float4)));
  {
    // a whole number of 2 MiB pages: the accumulator then owns its driver allocation, so that the CUDA IPC handle exported
    // for the peer-memory read-out (rt2_accum_ipc_handle) names exactly this buffer (small cudaMalloc blocks are
    // sub-allocated from shared 2 MiB pages, and an IPC handle always names the whole page)
    const size_t page = 2ull << 20;
    RT2_CUDA(cudaMalloc(&m.accum, (P * sizeof(float4) + page - 1) / page * page));
  }
  if (cfg_.flags & RT2_FLAG_MOMENTS) RT2_CUDA(cudaMalloc(&m.accum_sq, P * sizeof(float4)));
  RT2_CUDA(cudaMalloc(&m.mean_rgb, P * 3 * sizeof(float)));
  RT2_CUDA(cudaMalloc(&m.rgba8, P * sizeof(uchar4)));
#ifdef RT2_WITH_RAY_SORT
  if (m.sort_enabled) {
    for (int i = 0; i < 2; i++) {
      RT2_CUDA(cudaMalloc(&m.sort_keys[i], N * sizeof(uint32_t)));
      RT2_CUDA(cudaMalloc(&m.sort_vals[i], N * sizeof(uint32_t)));
    }
    RT2_CUDA(cudaMalloc(&m.sort_hist, static_cast<size_t>(m.grid_sort) * kSortBins * sizeof(uint32_t)));
    RT2_CUDA(cudaMalloc(&m.sort_bin_base, kSortBins * sizeof(uint32_t)));
  }
#endif
#ifdef RT2_DEBUG_CHECKS
  // poison: a kernel that reads wavefront state nobody wrote produces NaNs (bit pattern 0xFF..) instead of plausible zeros
  for (int i = 0; i < 2; i++) {
    RT2_CUDA(cudaMemset(m.ray_o[i], 0xFF, N * sizeof(float4)));
    RT2_CUDA(cudaMemset(m.ray_d[i], 0xFF, N * sizeof(float4)));
    RT2_CUDA(cudaMemset(m.state[i], 0xFF, N * sizeof(float4)));
  }
  RT2_CUDA(cudaMemset(m.hit0, 0xFF, N * sizeof(float4)));
  RT2_CUDA(cudaMemset(m.hit1, 0xFF, N * sizeof(float4)));
  RT2_CUDA(cudaMemset(m.trav, 0xFF, N * sizeof(uint4)));
  RT2_CUDA(cudaMemset(m.bins.base, 0xFF, N * kNumBins * sizeof(uint32_t)));
  RT2_CUDA(cudaMemset(m.mean_rgb, 0xFF, P * 3 * sizeof(float)));
#endif
  if (m.split_mode) return AllocSplitState();
  return RT2_OK;
}

// Ray-queue sizes of the most recent wavefront batch, one per bounce (tools: per-bounce analysis, ncu bytes-per-ray).
int Renderer::QueueSizes(uint32_t* out, uint32_t max_bounces, uint32_t* n_bounces) {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  if (!state_ok_) return NoState();
  int rc = Synchronize();
  if (rc != RT2_OK) return rc;
  const uint32_t depth = static_cast<uint32_t>(cfg_.max_depth);
  std::vector<uint32_t> ctr(static_cast<size_t>(depth) * kCounterStride);
  RT2_CUDA(cudaMemcpy(ctr.data(), m.counters, ctr.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  *n_bounces = depth;
  for (uint32_t b = 0; b < depth && b < max_bounces; b++) out[b] = ctr[static_cast<size_t>(b) * kCounterStride];
  return RT2_OK;
}

// DEBUG_CHECKS builds: the violation counters of the device-side self checks (rt_trace.cuh); all zero otherwise.
int Renderer::DebugCounters(uint64_t* out, int* enabled) {
  for (int k = 0; k < 16; k++) out[k] = 0;
#ifdef RT2_DEBUG_CHECKS
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  RT2_CUDA(cudaStreamSynchronize(m.stream));
  unsigned long long v[kChkCount];
  RT2_CUDA(cudaMemcpyFromSymbol(v, g_rt2_violations, sizeof(v)));
  for (int k = 0; k < kChkCount && k < 16; k++) out[k] = v[k];
  *enabled = 1;
#else
  *enabled = 0;
#endif
  return RT2_OK;
}

int Renderer::Reset() {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  if (!state_ok_) return NoState();
  pending_frames_ = 0;  // frames requested but not yet traced are dropped with the accumulators (RayTracer.cpp:49-53)
  const size_t P = static_cast<size_t>(width_) * height_;
  RT2_CUDA(cudaMemsetAsync(m.accum, 0, P * sizeof(float4), m.stream));
  if (m.accum_sq) RT2_CUDA(cudaMemsetAsync(m.accum_sq, 0, P * sizeof(float4), m.stream));
  RT2_CUDA(cudaMemsetAsync(m.totals, 0, 8 * sizeof(unsigned long long), m.stream));
  frame_idx_ = 0;
  gpu_ms_total_ = 0;
  for (double& v : prof_ms_) v = 0;
  return RT2_OK;
}

// One extend stage = closest surface of every queued ray, by whichever traversal the scene uses.
struct ExtendArgs {
  const uint32_t* n_ptr;   // device-resident ray count, or nullptr: n_fixed
  uint32_t n_fixed;
  uint32_t* cursor;        // fetch cursor of the ray queue
  uint32_t* entry_count;   // instance split: entry counter and fetch cursor of this stage
  uint32_t* entry_cursor;
  const float4* ray_o;
  const float4* ray_d;
  float tmin, tmax;
  const uint32_t* order;
  uint32_t sort_min_rays;
  uint4* trav;
  SplitIO io;
  int grid;                // persistent grid of the BVH walks
  int grid_flat;           // grid of the flat kernel
};

template <class M, bool kCount>
static void LaunchExtendT(Renderer::Impl& m, const ExtendArgs& a, uint64_t* launches, cudaEvent_t mid) {
  if (m.wide_mode) {
    k_traverse_wide<M, kCount><<<a.grid, kBlock, 0, m.stream>>>(m.ds, m.wide, a.n_ptr, a.n_fixed, a.cursor, a.ray_o, a.ray_d, a.tmin, a.tmax,
                                                                a.trav, m.totals, m.trav_max_steps);
    (*launches)++;
  } else if (m.flat_mode) {
    k_traverse_flat<M><<<a.grid_flat, kBlock, 0, m.stream>>>(m.ds, a.n_ptr, a.n_fixed, a.ray_o, a.ray_d, a.tmin, a.tmax, a.trav);
    (*launches)++;
  } else if (m.unified_mode) {
    if (m.quant_nodes)
      k_traverse<M, kCount, kTravUnified, true><<<a.grid, kBlock, 0, m.stream>>>(m.ds, a.n_ptr, a.n_fixed, a.cursor, a.ray_o, a.ray_d, a.tmin,
                                                                                    a.tmax, a.order, a.sort_min_rays, a.trav, SplitIO{},
                                                                                    m.totals, m.trav_max_steps, m.trav_fetch_threshold);
    else
      k_traverse<M, kCount, kTravUnified><<<a.grid, kBlock, 0, m.stream>>>(m.ds, a.n_ptr, a.n_fixed, a.cursor, a.ray_o, a.ray_d, a.tmin, a.tmax,
                                                                              a.order, a.sort_min_rays, a.trav, SplitIO{}, m.totals,
                                                                              m.trav_max_steps, m.trav_fetch_threshold);
    (*launches)++;
  } else if (m.split_mode) {
    SplitIO io = a.io;
    io.entry_count = a.entry_count;
    if (io.capacity == 0) io.capacity = a.n_fixed * m.ds.n_hoisted;  // rt2_intersect sizes its own queue
    k_traverse<M, kCount, kTravWorld><<<a.grid, kBlock, 0, m.stream>>>(m.ds, a.n_ptr, a.n_fixed, a.cursor, a.ray_o, a.ray_d, a.tmin, a.tmax,
                                                                          a.order, a.sort_min_rays, a.trav, io, m.totals, m.trav_max_steps,
                                                                          m.trav_fetch_threshold);
    if (mid) cudaEventRecord(mid, m.stream);  // profiling: world pass | instance pass
    k_traverse<M, kCount, kTravInst><<<a.grid, kBlock, 0, m.stream>>>(m.ds, a.entry_count, 0u, a.entry_cursor, a.ray_o, a.ray_d, a.tmin,
                                                                         a.tmax, nullptr, 0u, a.trav, io, m.totals, m.trav_max_steps,
                                                                         m.trav_fetch_threshold);
    *launches += 2;
  } else {
    if (m.quant_nodes)
      k_traverse<M, kCount, kTravInline, true><<<a.grid, kBlock, 0, m.stream>>>(m.ds, a.n_ptr, a.n_fixed, a.cursor, a.ray_o, a.ray_d, a.tmin,
                                                                                   a.tmax, a.order, a.sort_min_rays, a.trav, SplitIO{},
                                                                                   m.totals, m.trav_max_steps, m.trav_fetch_threshold);
    else
      k_traverse<M, kCount, kTravInline><<<a.grid, kBlock, 0, m.stream>>>(m.ds, a.n_ptr, a.n_fixed, a.cursor, a.ray_o, a.ray_d, a.tmin, a.tmax,
                                                                             a.order, a.sort_min_rays, a.trav, SplitIO{}, m.totals,
                                                                             m.trav_max_steps, m.trav_fetch_threshold);
    (*launches)++;
  }
}

static void LaunchExtend(Renderer::Impl& m, const ExtendArgs& a, bool exact, bool count, uint64_t* launches, cudaEvent_t mid = nullptr) {
  if (exact) {
    if (count) LaunchExtendT<ExactMath, true>(m, a, launches, mid);
    else LaunchExtendT<ExactMath, false>(m, a, launches, mid);
  } else {
    if (count) LaunchExtendT<FastMath, true>(m, a, launches, mid);
    else LaunchExtendT<FastMath, false>(m, a, launches, mid);
  }
}

template <int kType, int kBin>
static void LaunchScatter(const Renderer::Impl& m, const FrameParams& fp, uint32_t bounce, uint32_t* ctr, int in, int out,
                          uint32_t* sort_keys, const uint32_t* inline_count) {
  k_shade_scatter<kType, kBin><<<m.grid_stream, kBlock, 0, m.stream>>>(m.ds, fp, bounce, ctr, m.bins.q(kBin), m.ray_o[in], m.ray_d[in],
                                                                       m.state[in], m.hit0, m.hit1, m.ray_o[out], m.ray_d[out],
                                                                       m.state[out], sort_keys, inline_count);
}

int Renderer::RenderBatch(uint32_t n_frames) {
  Impl& m = *impl_;
  const uint32_t P = static_cast<uint32_t>(width_) * height_;
  const uint32_t n_slots = P * n_frames;
  FrameParams fp;
  fp.cam = camera_;
  fp.width = width_;
  fp.height = height_;
  fp.pixels = P;
  // Camera::Update: sqrt_samples_per_pix_ = int(sqrt(samples_per_pixel)), recip = 1.0 / sqrt (Camera.hpp:45-47)
  int sq = static_cast<int>(std::sqrt(static_cast<double>(cfg_.samples_per_pixel)));
  if (sq < 1) sq = 1;
  fp.sqrt_spp = sq;
  fp.recip_sqrt_spp = static_cast<float>(1.0 / sq);
  fp.frame_stride = cfg_.frame_stride;
  fp.frame_base = static_cast<uint32_t>(cfg_.frame_offset + frame_idx_ * cfg_.frame_stride);
  fp.seed_lo = static_cast<uint32_t>(cfg_.seed);
  fp.seed_hi = static_cast<uint32_t>(cfg_.seed >> 32);
  fp.pixels_magic = P > 1 ? (~0ull / P + 1ull) : 0ull;  // ceil(2^64 / P), also when P is a power of two
  const bool exact = !(cfg_.flags & RT2_FLAG_FAST_MATH);
  const uint32_t max_depth = static_cast<uint32_t>(cfg_.max_depth);

  auto prof = [&](int kind, bool record = true) -> cudaEvent_t {
    if (!profiling_) return nullptr;
    cudaEvent_t e;
    if (prof_used_ < m.prof_events.size()) {
      e = m.prof_events[prof_used_];
    } else {
      cudaEventCreate(&e);
      m.prof_events.push_back(e);
      m.prof_kind.push_back(0);
    }
    m.prof_kind[prof_used_] = kind;
    prof_used_++;
    if (record) cudaEventRecord(e, m.stream);
    return e;
  };
  prof_used_ = 0;

  RT2_CUDA(cudaMemsetAsync(m.counters, 0, static_cast<size_t>(max_depth + 1) * kCounterStride * sizeof(uint32_t), m.stream));
  RT2_CUDA(cudaMemsetAsync(m.radiance, 0, static_cast<size_t>(n_slots) * sizeof(float4), m.stream));
  prof(0);
  k_generate<<<m.grid_stream, kBlock, 0, m.stream>>>(fp, n_slots, m.ray_o[0], m.ray_d[0], m.state[0], m.counters);
  launches_++;
  int in = 0;
  for (uint32_t b = 0; b < max_depth; b++) {
    uint32_t* ctr = m.counters + static_cast<size_t>(b) * kCounterStride;
    uint32_t* next = m.counters + static_cast<size_t>(b + 1) * kCounterStride;
    const int out = in ^ 1;
    // coherence order of this bounce's queue (keys were written by the scatter kernels of the previous bounce)
    const bool sorted = m.sort_enabled && !m.flat_mode && b >= 1 && b <= m.sort_max_bounce;  // the flat extend kernel has no use for an order
    const bool sort_next = m.sort_enabled && !m.flat_mode && b + 1 <= m.sort_max_bounce;
    const uint32_t* order = nullptr;
#ifdef RT2_WITH_RAY_SORT
    if (sorted) {
      prof(5);
      const uint32_t smin = m.sort_min_rays;
      k_sort_hist<<<m.grid_sort, kSortBlock, 0, m.stream>>>(ctr, smin, m.sort_keys[0], 0, m.sort_hist);
      k_sort_scan<<<1, kSortBins, 0, m.stream>>>(ctr, smin, m.grid_sort, m.sort_hist, m.sort_bin_base);
      k_sort_scatter<<<m.grid_sort, kSortBlock, 0, m.stream>>>(ctr, smin, m.sort_keys[0], nullptr, 0, m.sort_hist, m.sort_bin_base,
                                                               m.sort_keys[1], m.sort_vals[1]);
      k_sort_hist<<<m.grid_sort, kSortBlock, 0, m.stream>>>(ctr, smin, m.sort_keys[1], kSortDigitBits, m.sort_hist);
      k_sort_scan<<<1, kSortBins, 0, m.stream>>>(ctr, smin, m.grid_sort, m.sort_hist, m.sort_bin_base);
      k_sort_scatter<<<m.grid_sort, kSortBlock, 0, m.stream>>>(ctr, smin, m.sort_keys[1], m.sort_vals[1], kSortDigitBits, m.sort_hist,
                                                               m.sort_bin_base, nullptr, m.sort_vals[0]);
      launches_ += 6;
      order = m.sort_vals[0];
    }
#else
    (void)sorted;
#endif
    prof(1);
    {
      ExtendArgs ea{};
      ea.n_ptr = ctr;
      ea.cursor = ctr + 7;
      ea.entry_count = ctr + 8;
      ea.entry_cursor = ctr + 9;
      ea.ray_o = m.ray_o[in];
      ea.ray_d = m.ray_d[in];
      ea.tmin = 0.001f;  // Interval{0.001, kInfinity} (RayTracer.cpp:25)
      ea.tmax = kFltMax;
      ea.order = order;
      ea.sort_min_rays = m.sort_min_rays;
      ea.trav = m.trav;
      ea.io = m.split;
      ea.grid = m.grid_extend;
      ea.grid_flat = m.grid_stream;
      // profiling: the event between the two passes of the instance split is recorded by LaunchExtend (kind 6 = instance pass)
      cudaEvent_t mid = (m.split_mode && !m.wide_mode && !m.flat_mode) ? prof(6, false) : nullptr;
      LaunchExtend(m, ea, exact, profiling_, &launches_, mid);
    }
    prof(3);
    const bool last = (b + 1 == max_depth);  // RayColor(depth <= 0) returns black: nothing to scatter into
    uint32_t* keys = (sort_next && !last) ? m.sort_keys[0] : nullptr;
    const SplitIO io = m.split;  // all-null unless the instance split is active (then load_closest merges the entries)
    if (m.fused) {
      // finish + inline shade; only noise-textured materials go through the bins
      const bool media_free = exact && (m.ds.n_media == 0 || m.split_media);  // no media, or sampled by k_media first
      if (media_free && m.split_media) {
        k_media<ExactMath><<<m.grid_stream, kBlock, 0, m.stream>>>(m.ds, fp, b, ctr, m.ray_o[in], m.ray_d[in], m.state[in], m.trav, io);
        launches_++;
      }
#define RT2_FS_ARGS                                                                                                                 \
  m.ds, fp, b, last ? 0 : 1, ctr, next, m.ray_o[in], m.ray_d[in], m.state[in], m.trav, io, m.hit0, m.hit1, m.bins, m.ray_o[out], \
      m.ray_d[out], m.state[out], keys, m.radiance
      const bool images = m.ds.n_images != 0;
      if (media_free && !images) {
        k_finish_shade<ExactMath, 4, 0, false><<<m.grid_stream, kBlock, 0, m.stream>>>(RT2_FS_ARGS);
      } else if (media_free) {
        k_finish_shade<ExactMath, 4, 0, true><<<m.grid_stream, kBlock, 0, m.stream>>>(RT2_FS_ARGS);
      } else if (exact && m.simple_media && !images) {
        k_finish_shade<ExactMath, 4, 1, false><<<m.grid_stream, kBlock, 0, m.stream>>>(RT2_FS_ARGS);
      } else if (exact) {
        k_finish_shade<ExactMath><<<m.grid_stream, kBlock, 0, m.stream>>>(RT2_FS_ARGS);
      } else {
        k_finish_shade<FastMath><<<m.grid_stream, kBlock, 0, m.stream>>>(RT2_FS_ARGS);
      }
#undef RT2_FS_ARGS
      launches_++;
      prof(2);
      if (!last) {
        if (m.deferred_present[2]) LaunchScatter<RT2_MAT_TEXTURE, 2>(m, fp, b, ctr, in, out, keys, next), launches_++;
        if (m.deferred_present[5]) LaunchScatter<RT2_MAT_ISOTROPIC, 5>(m, fp, b, ctr, in, out, keys, next), launches_++;
      }
      if (m.deferred_present[0] || ((m.deferred_present[2] || m.deferred_present[5]) && !last)) {
        k_shade_terminal<<<m.grid_stream, kBlock, 0, m.stream>>>(m.ds, ctr, last ? nullptr : next, m.bins.q(0), m.state[in], m.hit0, m.hit1,
                                                                 m.radiance, 1);
        launches_++;
      }
    } else {
      if (exact) {
        k_finish_hit<ExactMath><<<m.grid_stream, kBlock, 0, m.stream>>>(m.ds, fp, b, ctr, m.ray_o[in], m.ray_d[in], m.state[in], m.trav, io,
                                                                        m.hit0, m.hit1, m.bins);
      } else {
        k_finish_hit<FastMath><<<m.grid_stream, kBlock, 0, m.stream>>>(m.ds, fp, b, ctr, m.ray_o[in], m.ray_d[in], m.state[in], m.trav, io,
                                                                       m.hit0, m.hit1, m.bins);
      }
      launches_++;
      prof(2);
      k_shade_terminal<<<m.grid_stream, kBlock, 0, m.stream>>>(m.ds, ctr, last ? nullptr : next, m.bins.q(0), m.state[in], m.hit0, m.hit1,
                                                               m.radiance, 0);
      launches_++;
      if (!last) {
        if (m.bin_present[1]) LaunchScatter<RT2_MAT_LAMBERTIAN, 1>(m, fp, b, ctr, in, out, keys, nullptr), launches_++;
        if (m.bin_present[2]) LaunchScatter<RT2_MAT_TEXTURE, 2>(m, fp, b, ctr, in, out, keys, nullptr), launches_++;
        if (m.bin_present[3]) LaunchScatter<RT2_MAT_METAL, 3>(m, fp, b, ctr, in, out, keys, nullptr), launches_++;
        if (m.bin_present[4]) LaunchScatter<RT2_MAT_DIELECTRIC, 4>(m, fp, b, ctr, in, out, keys, nullptr), launches_++;
        if (m.bin_present[5]) LaunchScatter<RT2_MAT_ISOTROPIC, 5>(m, fp, b, ctr, in, out, keys, nullptr), launches_++;
      }
    }
    in = out;
  }
  prof(0);
  k_accumulate<<<m.grid_stream, kBlock, 0, m.stream>>>(P, n_frames, m.radiance, m.accum, m.accum_sq);
  k_batch_stats<<<1, 32, 0, m.stream>>>(m.counters, max_depth, m.totals);
  launches_ += 2;
  prof(4);
  RT2_CUDA(cudaGetLastError());
  if (profiling_) {
    RT2_CUDA(cudaStreamSynchronize(m.stream));
    for (size_t i = 0; i + 1 < prof_used_; i++) {
      float ms = 0;
      cudaEventElapsedTime(&ms, m.prof_events[i], m.prof_events[i + 1]);
      int kind = m.prof_kind[i];
      if (kind >= 0 && kind < 7 && kind != 4) prof_ms_[kind] += ms;
    }
  }
  return RT2_OK;
}

int Renderer::NoState() {
  err_ = "the renderer has no frame buffers (the last rt2_resize failed): resize it first";
  return RT2_ERR_STATE;
}

// n x RayTracer::Update (RayTracer.cpp:55-70).  The reference's app calls Update once per sample (App.cpp:244-246); tracing one
// 600 x 600 frame per call would launch ~150 kernels over 360 k paths, so frames are collected until a wavefront batch is
// full and traced then — or at the next call that needs them (Flush).  FrameIdx() counts traced + pending frames, the
// stratum and Philox counters of a frame depend only on its index, so the image is the same as with eager tracing.
int Renderer::Update(uint32_t n_frames) {
  if (!state_ok_) return NoState();
  pending_frames_ += n_frames;
  while (pending_frames_ >= static_cast<uint64_t>(frames_per_batch_)) {
    int rc = TraceFrames(static_cast<uint32_t>(frames_per_batch_));
    if (rc != RT2_OK) return rc;
  }
  return RT2_OK;
}

int Renderer::Flush() {
  if (!state_ok_) return pending_frames_ ? NoState() : RT2_OK;
  while (pending_frames_ > 0) {
    const uint64_t f = pending_frames_ < static_cast<uint64_t>(frames_per_batch_) ? pending_frames_ : static_cast<uint64_t>(frames_per_batch_);
    int rc = TraceFrames(static_cast<uint32_t>(f));
    if (rc != RT2_OK) return rc;
  }
  return RT2_OK;
}

int Renderer::TraceFrames(uint32_t f) {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  // one CUDA-event pair per batch; folded into gpu_ms_total at the next synchronisation
  if (m.timing_used + 2 > m.timing_events.size()) {
    for (int k = 0; k < 2; k++) {
      cudaEvent_t e;
      RT2_CUDA(cudaEventCreate(&e));
      m.timing_events.push_back(e);
    }
  }
  cudaEvent_t ev_a = m.timing_events[m.timing_used], ev_b = m.timing_events[m.timing_used + 1];
  m.timing_used += 2;
  RT2_CUDA(cudaEventRecord(ev_a, m.stream));
  int rc = RenderBatch(f);
  if (rc != RT2_OK) return rc;
  frame_idx_ += f;
  pending_frames_ -= f;
  RT2_CUDA(cudaEventRecord(ev_b, m.stream));
  timing_pending_ = true;
  return RT2_OK;
}

int Renderer::Synchronize() {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  int rc = Flush();
  if (rc != RT2_OK) return rc;
  RT2_CUDA(cudaStreamSynchronize(m.stream));
  if (timing_pending_) {
    for (size_t i = 0; i + 1 < m.timing_used; i += 2) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, m.timing_events[i], m.timing_events[i + 1]) == cudaSuccess) gpu_ms_total_ += ms;
    }
    m.timing_used = 0;
    timing_pending_ = false;
  }
  return RT2_OK;
}

// A traversal that ran out of stack dropped a sub-tree: the image may miss hits.  Reported by every read-out instead of
// returning a silently wrong image (never expected: the builders bound the depth of every tree).
int Renderer::CheckOverflow() {
  Impl& m = *impl_;
  unsigned long long ov = 0;
  RT2_CUDA(cudaMemcpy(&ov, m.totals + 6, sizeof(ov), cudaMemcpyDeviceToHost));
  if (ov != 0) {
    err_ = "BVH traversal stack overflow in " + std::to_string(ov) + " warp(s): a tree is deeper than the " + std::to_string(kStackSize) +
           "-entry stack, the image is incomplete";
    return RT2_ERR_STATE;
  }
  return RT2_OK;
}

int Renderer::ReadMean(float* dst) {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  if (!state_ok_) return NoState();
  int rcf = Flush();
  if (rcf != RT2_OK) return rcf;
  const uint32_t P = static_cast<uint32_t>(width_) * height_;
  // accum / frame_idx_ — with no frames the reference divides by zero (NaN); we do the same
  k_resolve<<<m.grid_stream, kBlock, 0, m.stream>>>(P, static_cast<float>(frame_idx_), m.accum, m.mean_rgb, nullptr);
  launches_++;
  const size_t bytes = static_cast<size_t>(P) * 3 * sizeof(float);
  int rc = EnsureStage(m, bytes, &err_);
  if (rc != RT2_OK) return rc;
  if (m.h_stage_cap >= bytes) {  // D2H into page-locked memory, then a host copy into the caller's buffer
    RT2_CUDA(cudaMemcpyAsync(m.h_stage, m.mean_rgb, bytes, cudaMemcpyDeviceToHost, m.stream));
    rc = Synchronize();
    if (rc == RT2_OK) rc = CheckOverflow();
    if (rc == RT2_OK) std::memcpy(dst, m.h_stage, bytes);
    return rc;
  }
  RT2_CUDA(cudaMemcpyAsync(dst, m.mean_rgb, bytes, cudaMemcpyDeviceToHost, m.stream));
  rc = Synchronize();
  return rc == RT2_OK ? CheckOverflow() : rc;
}

int Renderer::ReadRGBA8(uint8_t* dst) {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  if (!state_ok_) return NoState();
  int rcf = Flush();
  if (rcf != RT2_OK) return rcf;
  const uint32_t P = static_cast<uint32_t>(width_) * height_;
  k_resolve<<<m.grid_stream, kBlock, 0, m.stream>>>(P, static_cast<float>(frame_idx_), m.accum, nullptr, m.rgba8);
  launches_++;
  const size_t bytes = static_cast<size_t>(P) * 4;
  int rc = EnsureStage(m, bytes, &err_);
  if (rc != RT2_OK) return rc;
  if (m.h_stage_cap >= bytes) {
    RT2_CUDA(cudaMemcpyAsync(m.h_stage, m.rgba8, bytes, cudaMemcpyDeviceToHost, m.stream));
    rc = Synchronize();
    if (rc == RT2_OK) rc = CheckOverflow();
    if (rc == RT2_OK) std::memcpy(dst, m.h_stage, bytes);
    return rc;
  }
  RT2_CUDA(cudaMemcpyAsync(dst, m.rgba8, bytes, cudaMemcpyDeviceToHost, m.stream));
  rc = Synchronize();
  return rc == RT2_OK ? CheckOverflow() : rc;
}

int Renderer::ReadAccum(float* sum, float* sumsq) {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  if (!state_ok_) return NoState();
  int rc = Synchronize();
  if (rc == RT2_OK) rc = CheckOverflow();
  if (rc != RT2_OK) return rc;
  const size_t P = static_cast<size_t>(width_) * height_;
  std::vector<float4> tmp(P);
  auto fetch = [&](const float4* src, float* dst) -> int {
    RT2_CUDA(cudaMemcpy(tmp.data(), src, P * sizeof(float4), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < P; i++) {
      dst[3 * i + 0] = tmp[i].x;
      dst[3 * i + 1] = tmp[i].y;
      dst[3 * i + 2] = tmp[i].z;
    }
    return RT2_OK;
  };
  if (sum) {
    rc = fetch(m.accum, sum);
    if (rc != RT2_OK) return rc;
  }
  if (sumsq) {
    if (!m.accum_sq) {
      err_ = "renderer was created without RT2_FLAG_MOMENTS";
      return RT2_ERR_STATE;
    }
    rc = fetch(m.accum_sq, sumsq);
    if (rc != RT2_OK) return rc;
  }
  return RT2_OK;
}

// Restores the accumulators (checkpoint / resume): sum / sumsq as W*H*3 floats, `frames` = samples they hold.
int Renderer::WriteAccum(const float* sum, const float* sumsq, uint64_t frames) {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  if (!state_ok_) return NoState();
  pending_frames_ = 0;  // the restored accumulators replace whatever was requested before
  int rc = Synchronize();
  if (rc != RT2_OK) return rc;
  const size_t P = static_cast<size_t>(width_) * height_;
  std::vector<float4> tmp(P);
  auto put = [&](const float* src, float4* dst) -> int {
    for (size_t i = 0; i < P; i++) tmp[i] = make_float4(src[3 * i + 0], src[3 * i + 1], src[3 * i + 2], 0.0f);
    RT2_CUDA(cudaMemcpy(dst, tmp.data(), P * sizeof(float4), cudaMemcpyHostToDevice));
    return RT2_OK;
  };
  if (!sum) {
    err_ = "null sum";
    return RT2_ERR_INVALID_ARG;
  }
  rc = put(sum, m.accum);
  if (rc != RT2_OK) return rc;
  if (sumsq) {
    if (!m.accum_sq) {
      err_ = "renderer was created without RT2_FLAG_MOMENTS";
      return RT2_ERR_STATE;
    }
    rc = put(sumsq, m.accum_sq);
    if (rc != RT2_OK) return rc;
  }
  frame_idx_ = frames;
  return RT2_OK;
}

// handle layout (RT2_IPC_HANDLE_BYTES = 80): cudaIpcMemHandle_t (64) | byte offset of the accumulator inside the exported
// 2 MiB-aligned allocation (8) | reserved (8).  cudaIpcGetMemHandle names the whole driver allocation a pointer lives in
// and cudaIpcOpenMemHandle returns that allocation's base in the peer (measured: with a sub-allocated 1 MiB accumulator
// the peer read a neighbouring buffer).
int Renderer::AccumIpcHandle(uint8_t* handle) {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "layout of the exported handle");
  cudaIpcMemHandle_t h;
  RT2_CUDA(cudaIpcGetMemHandle(&h, m.accum));
  // the accumulator is allocated in whole 2 MiB pages (Resize), so it starts its driver allocation; the offset field is kept
  // for the page-relative position should that ever change
  const unsigned long long offset = reinterpret_cast<unsigned long long>(m.accum) & ((2ull << 20) - 1ull);
  std::memset(handle, 0, RT2_IPC_HANDLE_BYTES);
  std::memcpy(handle, &h, 64);
  std::memcpy(handle + 64, &offset, 8);
  return RT2_OK;
}

int Renderer::ResolvePeers(const uint8_t* handles, uint32_t n_ranks, uint32_t self_rank, uint64_t total_frames, float* dst_mean,
                           uint8_t* dst_rgba8) {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  if (!state_ok_) return NoState();
  if (n_ranks < 1 || n_ranks > static_cast<uint32_t>(kMaxPeers) || self_rank >= n_ranks || (n_ranks > 1 && !handles)) {
    err_ = "rt2_resolve_peers: bad rank arguments";
    return RT2_ERR_INVALID_ARG;
  }
  int rc = Flush();
  if (rc != RT2_OK) return rc;
  const uint32_t P = static_cast<uint32_t>(width_) * height_;
  const size_t b_mean = static_cast<size_t>(P) * 3 * sizeof(float), b_rgba = static_cast<size_t>(P) * 4;
  // Every allocation this call needs is made BEFORE the peers are mapped, and the mappings are closed before returning:
  // on the B200 boxes a cudaMallocHost issued while a lazily-enabled IPC mapping was open was handed the mapping's own
  // virtual address (measured: the second read-out then read the staging buffer instead of the peer), so no mapping
  // outlives the call.
  rc = EnsureStage(m, b_mean + b_rgba, &err_);
  if (rc != RT2_OK) return rc;
  const void* ptrs[kMaxPeers];
  void* opened[kMaxPeers];
  int n_opened = 0;
  auto close_all = [&]() {
    for (int k = 0; k < n_opened; k++) cudaIpcCloseMemHandle(opened[k]);
    n_opened = 0;
  };
  for (uint32_t r = 0; r < n_ranks; r++) {
    if (r == self_rank) {
      ptrs[r] = m.accum;
      continue;
    }
    const uint8_t* hr = handles + static_cast<size_t>(RT2_IPC_HANDLE_BYTES) * r;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, hr, 64);
    unsigned long long offset = 0;
    std::memcpy(&offset, hr + 64, 8);
    void* mapped = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      close_all();
      err_ = std::string("cudaIpcOpenMemHandle failed: ") + cudaGetErrorString(e);
      return RT2_ERR_CUDA;
    }
    opened[n_opened++] = mapped;
    ptrs[r] = static_cast<const char*>(mapped) + offset;
  }
  rc = ResolvePointers(ptrs, n_ranks, total_frames, dst_mean, dst_rgba8);
  close_all();
  return rc;
}

// Sum of `n` accumulators (W*H float4 each, any device this one can address: own HBM, peer HBM over NVLink) in the order
// given, divided by total_frames: mean and / or RGBA8 preview, one kernel on this renderer's stream, then the read-back.
int Renderer::ResolvePointers(const void* const* accums, uint32_t n, uint64_t total_frames, float* dst_mean, uint8_t* dst_rgba8) {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  if (!state_ok_) return NoState();
  if (n < 1 || n > static_cast<uint32_t>(kMaxPeers)) {
    err_ = "resolve: between 1 and " + std::to_string(kMaxPeers) + " accumulators";
    return RT2_ERR_INVALID_ARG;
  }
  const uint32_t P = static_cast<uint32_t>(width_) * height_;
  const size_t b_mean = static_cast<size_t>(P) * 3 * sizeof(float), b_rgba = static_cast<size_t>(P) * 4;
  int rc = EnsureStage(m, b_mean + b_rgba, &err_);
  if (rc != RT2_OK) return rc;
  const bool staged = m.h_stage_cap >= b_mean + b_rgba;
  PeerAccums pa{};
  for (uint32_t r = 0; r < n; r++) pa.p[r] = static_cast<const float4*>(accums[r]);
  k_resolve_peers<<<m.grid_stream, kBlock, 0, m.stream>>>(P, static_cast<float>(total_frames), pa, static_cast<int>(n),
                                                          dst_mean ? m.mean_rgb : nullptr, dst_rgba8 ? m.rgba8 : nullptr);
  launches_++;
  cudaError_t e = cudaSuccess;
  if (dst_mean) e = cudaMemcpyAsync(staged ? static_cast<void*>(m.h_stage) : dst_mean, m.mean_rgb, b_mean, cudaMemcpyDeviceToHost, m.stream);
  if (e == cudaSuccess && dst_rgba8)
    e = cudaMemcpyAsync(staged ? static_cast<void*>(m.h_stage + b_mean) : dst_rgba8, m.rgba8, b_rgba, cudaMemcpyDeviceToHost, m.stream);
  rc = (e == cudaSuccess) ? Synchronize() : RT2_ERR_CUDA;
  if (e != cudaSuccess) err_ = std::string("resolve: copy failed: ") + cudaGetErrorString(e);
  if (rc == RT2_OK) rc = CheckOverflow();
  if (rc == RT2_OK && staged) {
    if (dst_mean) std::memcpy(dst_mean, m.h_stage, b_mean);
    if (dst_rgba8) std::memcpy(dst_rgba8, m.h_stage + b_mean, b_rgba);
  }
  return rc;
}

// Multi-device plumbing (rt_multi.cpp): an event on this renderer's stream after everything queued so far, and a wait of
// this renderer's stream on another renderer's event.
int Renderer::RecordDone(void** event_out) {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  if (!m.ev_done) RT2_CUDA(cudaEventCreateWithFlags(&m.ev_done, cudaEventDisableTiming));
  RT2_CUDA(cudaEventRecord(m.ev_done, m.stream));
  *event_out = m.ev_done;
  return RT2_OK;
}
int Renderer::WaitFor(void* event) {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  RT2_CUDA(cudaStreamWaitEvent(m.stream, static_cast<cudaEvent_t>(event), 0));
  return RT2_OK;
}
void* Renderer::AccumPtr() { return impl_->accum; }
void* Renderer::AccumSqPtr() { return impl_->accum_sq; }

int Renderer::AccumDevicePtr(void** ptr, size_t* n_floats) {
  if (!state_ok_) return NoState();
  int rc = Flush();  // whoever asks for the accumulator wants every requested frame in it (queued on the stream)
  if (rc != RT2_OK) return rc;
  *ptr = impl_->accum;
  *n_floats = static_cast<size_t>(width_) * height_ * 4;
  return RT2_OK;
}

void* Renderer::Stream() { return impl_->stream; }

int Renderer::Intersect(const float* rays, size_t n, float tmin, float tmax, int skip_media, rt2_hit* out) {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  if (n == 0) return RT2_OK;
  if (n > 0x3FFFFFFFull) {
    err_ = "too many rays";
    return RT2_ERR_INVALID_ARG;
  }
  // split the interleaved input (origin xyz, time | direction xyz, pad) into the two arrays the traversal reads
  std::vector<float4> h_o(n), h_d(n);
  for (size_t i = 0; i < n; i++) {
    h_o[i] = make_float4(rays[8 * i + 0], rays[8 * i + 1], rays[8 * i + 2], rays[8 * i + 3]);
    h_d[i] = make_float4(rays[8 * i + 4], rays[8 * i + 5], rays[8 * i + 6], 0.0f);
  }
  float4* d_o = nullptr;
  float4* d_d = nullptr;
  uint4* d_trav = nullptr;
  uint32_t* d_ctr = nullptr;  // [0] ray cursor, [1] entry count, [2] entry cursor
  rt2_hit* d_out = nullptr;
  SplitIO io{};
  auto cleanup = [&]() {
    void* bufs[] = {d_o, d_d, d_trav, d_ctr, d_out, io.entries, io.entry_prim, io.inst_best};
    for (void* b : bufs)
      if (b) cudaFree(b);
  };
  cudaError_t e = cudaSuccess;
  if ((e = cudaMalloc(&d_o, n * sizeof(float4))) != cudaSuccess || (e = cudaMalloc(&d_d, n * sizeof(float4))) != cudaSuccess ||
      (e = cudaMalloc(&d_trav, n * sizeof(uint4))) != cudaSuccess || (e = cudaMalloc(&d_ctr, 4 * sizeof(uint32_t))) != cudaSuccess ||
      (e = cudaMalloc(&d_out, n * sizeof(rt2_hit))) != cudaSuccess) {
    cleanup();
    err_ = std::string("cudaMalloc failed: ") + cudaGetErrorString(e);
    return RT2_ERR_CUDA;
  }
  if (m.split_mode) {
    // the same two-pass extend the renderer runs: entry queue sized for these rays
    const size_t cap = n * m.ds.n_hoisted;
    if ((e = cudaMalloc(&io.entries, cap * sizeof(uint2))) != cudaSuccess || (e = cudaMalloc(&io.entry_prim, cap * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&io.inst_best, n * sizeof(unsigned long long))) != cudaSuccess) {
      cleanup();
      err_ = std::string("cudaMalloc failed: ") + cudaGetErrorString(e);
      return RT2_ERR_CUDA;
    }
  }
  cudaMemcpyAsync(d_o, h_o.data(), n * sizeof(float4), cudaMemcpyHostToDevice, m.stream);
  cudaMemcpyAsync(d_d, h_d.data(), n * sizeof(float4), cudaMemcpyHostToDevice, m.stream);
  cudaMemsetAsync(d_ctr, 0, 4 * sizeof(uint32_t), m.stream);
  const uint32_t n32 = static_cast<uint32_t>(n);
  const uint32_t grid = (n32 + kBlock - 1) / kBlock;
  const uint32_t seed_lo = static_cast<uint32_t>(cfg_.seed), seed_hi = static_cast<uint32_t>(cfg_.seed >> 32);
  const bool exact = !(cfg_.flags & RT2_FLAG_FAST_MATH);
  ExtendArgs ea{};
  ea.n_ptr = nullptr;
  ea.n_fixed = n32;
  ea.cursor = d_ctr;
  ea.entry_count = d_ctr + 1;
  ea.entry_cursor = d_ctr + 2;
  ea.ray_o = d_o;
  ea.ray_d = d_d;
  ea.tmin = tmin;
  ea.tmax = tmax;
  ea.trav = d_trav;
  ea.io = io;
  ea.grid = static_cast<int>(grid < static_cast<uint32_t>(m.grid_extend) ? grid : static_cast<uint32_t>(m.grid_extend));
  ea.grid_flat = static_cast<int>(grid);
  LaunchExtend(m, ea, exact, false, &launches_);
  if (exact) k_finish_intersect<ExactMath><<<grid, kBlock, 0, m.stream>>>(m.ds, d_o, d_d, d_trav, io, n32, tmin, skip_media, seed_lo, seed_hi, d_out);
  else k_finish_intersect<FastMath><<<grid, kBlock, 0, m.stream>>>(m.ds, d_o, d_d, d_trav, io, n32, tmin, skip_media, seed_lo, seed_hi, d_out);
  launches_++;
  cudaMemcpyAsync(out, d_out, n * sizeof(rt2_hit), cudaMemcpyDeviceToHost, m.stream);
  e = cudaStreamSynchronize(m.stream);
  cleanup();
  if (e != cudaSuccess) {
    err_ = std::string("rt2_intersect kernels failed: ") + cudaGetErrorString(e);
    return RT2_ERR_CUDA;
  }
  return CheckOverflow();
}

// Fixed-point parity hook for the device texture code: rgb[i] = textures[tex_idx].Value(u_i, v_i, p_i)
// (Texture.cpp:7-22, PerlinNoiseGen.cpp:52-88) evaluated by the same texture_value() the shade kernels call.
__global__ void __launch_bounds__(kBlock) k_texture_value(const DeviceScene S, uint32_t tex_idx, const float* __restrict__ pts,
                                                          const float* __restrict__ uv, uint32_t n, float* __restrict__ rgb) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const F3 p = {pts[3 * i + 0], pts[3 * i + 1], pts[3 * i + 2]};
  const F3 c = texture_value(S, tex_idx, p, uv ? uv[2 * i + 0] : 0.0f, uv ? uv[2 * i + 1] : 0.0f);
  rgb[3 * i + 0] = c.x;
  rgb[3 * i + 1] = c.y;
  rgb[3 * i + 2] = c.z;
}

int Renderer::TextureValue(uint32_t tex_idx, const float* points, const float* uv, size_t n, float* rgb) {
  Impl& m = *impl_;
  RT2_CUDA(cudaSetDevice(cfg_.device));
  if (n == 0) return RT2_OK;
  if (tex_idx >= n_textures_ || n > 0x3FFFFFFFull) {
    err_ = "rt2_texture_value: texture index out of range or too many points";
    return RT2_ERR_INVALID_ARG;
  }
  float *d_p = nullptr, *d_uv = nullptr, *d_rgb = nullptr;
  auto cleanup = [&]() {
    if (d_p) cudaFree(d_p);
    if (d_uv) cudaFree(d_uv);
    if (d_rgb) cudaFree(d_rgb);
  };
  cudaError_t e = cudaSuccess;
  if ((e = cudaMalloc(&d_p, n * 3 * sizeof(float))) != cudaSuccess || (e = cudaMalloc(&d_rgb, n * 3 * sizeof(float))) != cudaSuccess ||
      (uv && (e = cudaMalloc(&d_uv, n * 2 * sizeof(float))) != cudaSuccess)) {
    cleanup();
    err_ = std::string("cudaMalloc failed: ") + cudaGetErrorString(e);
    return RT2_ERR_CUDA;
  }
  cudaMemcpyAsync(d_p, points, n * 3 * sizeof(float), cudaMemcpyHostToDevice, m.stream);
  if (uv) cudaMemcpyAsync(d_uv, uv, n * 2 * sizeof(float), cudaMemcpyHostToDevice, m.stream);
  const uint32_t n32 = static_cast<uint32_t>(n);
  k_texture_value<<<(n32 + kBlock - 1) / kBlock, kBlock, 0, m.stream>>>(m.ds, tex_idx, d_p, d_uv, n32, d_rgb);
  launches_++;
  cudaMemcpyAsync(rgb, d_rgb, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, m.stream);
  e = cudaStreamSynchronize(m.stream);
  cleanup();
  if (e != cudaSuccess) {
    err_ = std::string("rt2_texture_value failed: ") + cudaGetErrorString(e);
    return RT2_ERR_CUDA;
  }
  return RT2_OK;
}

int Renderer::GetStats(rt2_stats* out) {
  Impl& m = *impl_;
  int rc = Synchronize();
  if (rc != RT2_OK) return rc;
  unsigned long long t[8];
  RT2_CUDA(cudaMemcpy(t, m.totals, sizeof(t), cudaMemcpyDeviceToHost));
  out->rays = t[0];
  out->paths = t[1];
  out->frames = frame_idx_;  // Synchronize() above traced every pending frame
  out->launches = launches_;
  out->gpu_ms_total = gpu_ms_total_;
  out->gpu_ms_other = prof_ms_[0];
  out->gpu_ms_extend = prof_ms_[1];
  out->gpu_ms_shade = prof_ms_[2];
  out->gpu_ms_finish = prof_ms_[3];
  out->gpu_ms_bvh_build = bvh_build_ms_;
  out->gpu_ms_sort = prof_ms_[5];
  out->gpu_ms_extend_inst = prof_ms_[6];
  out->box_pair_tests = t[2];
  out->sphere_tests = t[3];
  out->quad_tests = t[4];
  out->instance_visits = t[5];
  out->stack_overflows = t[6];
  out->max_stack_need = max_stack_need_;
  out->pending_frames = pending_frames_;
  out->n_gpus = 1;
  out->instance_split = m.split_mode ? 1u : 0u;
  out->compact_nodes = m.quant_nodes ? 1u : 0u;
  out->node_inflation = node_inflation_;
  out->instance_mode = m.flat_mode ? 4u : (m.unified_mode ? 3u : (m.split_mode ? 2u : (m.ds.n_instances ? 1u : 0u)));
  return RT2_OK;
}

// FP32 FMA peak of the device, measured: 8 independent FMA chains per thread, every SM fully occupied (the
// roofline the instruction-bound kernels are reported against must be measured, not nominal).
__global__ void __launch_bounds__(256) k_fma_peak(float* __restrict__ out, int iters, float a, float b) {
  float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
      x0 = fmaf(x0, a, b), x1 = fmaf(x1, a, b), x2 = fmaf(x2, a, b), x3 = fmaf(x3, a, b);
      x4 = fmaf(x4, a, b), x5 = fmaf(x5, a, b), x6 = fmaf(x6, a, b), x7 = fmaf(x7, a, b);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

int MeasureFp32Peak(int device, double* tflops, std::string* err) {
  auto fail = [&](const char* what, cudaError_t e) {
    *err = std::string(what) + " failed: " + cudaGetErrorString(e);
    return RT2_ERR_CUDA;
  };
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
  const int blocks = prop.multiProcessorCount * 8, iters = 4096;
  float* d = nullptr;
  if ((e = cudaMalloc(&d, static_cast<size_t>(blocks) * 256 * sizeof(float))) != cudaSuccess) return fail("cudaMalloc", e);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(a);
    k_fma_peak<<<blocks, 256>>>(d, iters, 0.999f, 0.001f);
    cudaEventRecord(b);
    if ((e = cudaEventSynchronize(b)) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double flops = 2.0 * 8 * 16 * static_cast<double>(iters) * blocks * 256;
    if (ms > 0 && flops / (ms * 1e-3) * 1e-12 > best) best = flops / (ms * 1e-3) * 1e-12;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  if (e != cudaSuccess) return fail("k_fma_peak", e);
  *tflops = best;
  return RT2_OK;
}

// L2 read bandwidth, measured: every block streams a 24 MiB buffer (L2-resident on a B200: 126 MB of L2, but far beyond
// the 148 x 256 KB of L1) with 16-byte loads, many times over; the first pass warms L2 and is not timed.
__global__ void __launch_bounds__(256) k_l2_read(const float4* __restrict__ buf, uint32_t n_vec, int passes, float* __restrict__ out) {
  float acc = 0.0f;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (int p = 0; p < passes; p++) {
    // rotate the start so consecutive passes of a block do not re-read what its own L1 still holds
    const uint32_t rot = static_cast<uint32_t>(p) * 0x9E3779B1u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
      const uint32_t j = (i + rot) % n_vec;
      float4 v;
      asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(buf + j));
      acc += v.x + v.y + v.z + v.w;
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

int MeasureL2Bandwidth(int device, double* gbs, std::string* err) {
  auto fail = [&](const char* what, cudaError_t e) {
    *err = std::string(what) + " failed: " + cudaGetErrorString(e);
    return RT2_ERR_CUDA;
  };
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail("cudaSetDevice", e);
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return fail("cudaGetDeviceProperties", e);
  const size_t bytes = 24ull << 20;
  const uint32_t n_vec = static_cast<uint32_t>(bytes / sizeof(float4));
  float4* d = nullptr;
  float* out = nullptr;
  if ((e = cudaMalloc(&d, bytes)) != cudaSuccess) return fail("cudaMalloc", e);
  if ((e = cudaMalloc(&out, sizeof(float))) != cudaSuccess) {
    cudaFree(d);
    return fail("cudaMalloc", e);
  }
  cudaMemset(d, 0, bytes);
  const int blocks = prop.multiProcessorCount * 8, passes = 16;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  double best = 0;
  k_l2_read<<<blocks, 256>>>(d, n_vec, 1, out);  // warm L2
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(a);
    k_l2_read<<<blocks, 256>>>(d, n_vec, passes, out);
    cudaEventRecord(b);
    if ((e = cudaEventSynchronize(b)) != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double g = static_cast<double>(bytes) * passes / (ms * 1e-3) * 1e-9;
    if (ms > 0 && g > best) best = g;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  cudaFree(out);
  if (e != cudaSuccess) return fail("k_l2_read", e);
  *gbs = best;
  return RT2_OK;
}

int DeviceCount() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

}  // namespace rt2
