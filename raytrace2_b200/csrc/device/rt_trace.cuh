// Closest-hit query on the flattened scene: replaces the reference's Hittable::Hit tree
// (BVHNode -> HittableList -> TransformedHittable -> Sphere / Quad / ConstantMedium).
//
// Contract: the result is the arg-min over all leaves of the RAW reported t, each leaf tested with the
// reference's own interval semantics (sphere: open, Sphere.cpp:21; quad: closed, Quad.cpp:27), instanced leaves report
// t in model units because TransformedHittable normalises the model-space direction but forwards ray_t unchanged
// (Transform.cpp:13-20,75-88).  Constant media are sampled after the surface traversal against the current best t,
// which is distributionally identical to the reference's in-traversal sampling — including the double
// draw for media that sit in a span-1 BVH leaf of the reference (Q2, BVH.cpp:18-20).
#pragma once
#include "../../../include/rt2.h"
#include "rt_math.cuh"

namespace rt2dev {

// Device-side self checks (compute-sanitizer is closed on this GPU pool, so the library carries its own): a DEBUG_CHECKS build
// (`make DEBUG_CHECKS=1` -> libraytrace2_b200_dbg.so) verifies every data-dependent index before it is used — node, primitive,
// material, texture, queue, entry and stack indices — counts violations per site in g_rt2_violations (never traps, so one run
// reports all of them) and poisons every wavefront buffer with NaN patterns at allocation (an uninitialised read then shows up
// as a NaN pixel).  tests/test_gpu_debug_checks.py runs every kernel family through it and asserts all counters are zero.
// The shipped library compiles the checks out.
enum {
  kChkNode = 0, kChkSphere, kChkQuad, kChkInstance, kChkInstLeaf, kChkStack, kChkQueue, kChkEntry, kChkMaterial, kChkTexture,
  kChkSlot, kChkBin, kChkMedium, kChkPrimRef, kChkNaN, kChkCount
};
#ifdef RT2_DEBUG_CHECKS
__device__ unsigned long long g_rt2_violations[kChkCount];
#define RT2_CHECK(cond, code)                                    \
  do {                                                           \
    if (!(cond)) atomicAdd(&g_rt2_violations[code], 1ull);       \
  } while (0)
#else
#define RT2_CHECK(cond, code) \
  do {                        \
  } while (0)
#endif

struct DeviceScene {
  const float4* __restrict__ spheres;    // 2 x float4 per sphere  {c0.xyz, r} {disp.xyz, mat}
  const float4* __restrict__ quads;      // 5 x float4 per quad    {n.xyz, d} {q.xyz, mat} {u} {v} {w}
  const float4* __restrict__ xforms;     // 6 x float4 per level   inv rows 0..2, model rows 0..2
  const uint4* __restrict__ instances;   // {chain_first, chain_len, blas_root, top_level}
  const uint4* __restrict__ media;       // 2 x uint4 per medium
  const float4* __restrict__ media_bounds;  // 2 x float4 per medium: padded box of the boundary in the medium's space
  const float4* __restrict__ materials;  // 2 x float4 per material
  const float4* __restrict__ textures;   // 3 x float4 per texture
  const rt2_perlin* __restrict__ perlin;
  const uint4* __restrict__ images;        // {texel_offset, width, height, 0} per image texture
  const float4* __restrict__ image_texels;  // linear RGBA, row 0 = top
  uint32_t n_images;
  // element counts of the buffers above (used by the DEBUG_CHECKS build only)
  uint32_t n_spheres, n_quads, n_materials, n_textures, n_node_pairs, n_prim_refs, n_inst_leaves;
  const uint32_t* __restrict__ prim_refs;
  const float4* __restrict__ nodes;  // 4 x float4 per node pair
  // compact node pairs (rt_qnodes.cu), nullptr when the scene keeps the float nodes: 2 x uint4 per pair, 15-bit boxes on one grid
  // per scene; a stored coordinate decodes to v in [1, 2) and means x = q_base + v * q_ext
  const uint4* __restrict__ qnodes;
  float q_base[3];
  float q_ext[3];
  uint32_t tlas_root;        // world TLAS with the instances as singleton leaves (kTravInline)
  uint32_t tlas_world_root;  // world TLAS over surfaces only (kTravWorld: the instances are hoisted out of the tree)
  uint32_t tlas_unified_root;  // world tree over surfaces + every instanced primitive as a world-space leaf (kTravUnified)
  const uint2* __restrict__ inst_leaves;  // kTravUnified: {primitive reference, instance index} of instanced leaf k
  uint32_t n_hoisted;        // > 0: instance split — all n_instances (<= kMaxHoistedInstances) are hoisted
  uint32_t n_media;
  uint32_t n_instances;
  float min_inv_scale;
  float background[3];
  // flat mode (tiny scenes, traverse_flat): every leaf primitive listed by space
  const uint32_t* __restrict__ flat_refs;     // world primitives, then the primitives of instance 0, 1, ...
  const uint32_t* __restrict__ flat_offsets;  // n_instances + 2 offsets into flat_refs
  const float4* __restrict__ inst_bounds;   // 2 x float4 per instance: conservative world-space box (flat mode, instance split)
  float sort_lo[3];     // ray-sort grid (rt_sort.cuh): robust scene bounds, 32 cells per axis
  float sort_scale[3];
};

constexpr float kFltMax = 3.402823466e+38f;  // kInfinity (Defs.hpp:17)
constexpr uint32_t kStackSentinel = 0x7FFFFFFFu;
constexpr uint32_t kLeafFlag = 0x80000000u;
// Traversal entries (the .w of a node's min corner on the device):
//   interior     child pair index                                              (bit 31 clear)
//   direct leaf  kLeafFlag | kLeafDirect | primitive reference (type << 28 | index)  — the usual case: one primitive per leaf
//   list leaf    kLeafFlag | (count - 1) << 26 | first                          — prim_refs[first .. first + count), count <= 16
constexpr uint32_t kLeafDirect = 0x40000000u;
__host__ __device__ inline uint32_t make_leaf_entry(uint32_t first, uint32_t count, uint32_t ref_if_single) {
  return count == 1u ? (kLeafFlag | kLeafDirect | ref_if_single) : (kLeafFlag | ((count - 1u) << 26) | first);
}
constexpr int kStackSize = 64;
// The 32-byte quantised node pairs are used when they grow the boxes' total surface area by at most this fraction.
constexpr float kQuantMaxInflation = 0.05f;

struct RaySpace {
  F3 o, d;
};

// Reciprocal direction for the slab test.  Exact zeros (and denormals) are replaced by +-1e-30 so that the fused form
// t = b * inv - o * inv never produces inf - inf; box tests only cull, the exact Hit() arithmetic never sees this value.
// Branch-free on purpose: a data-dependent branch here made ptxas wrap the traversal loop in extra convergence
// barriers and halved k_extend's throughput (measured on B200, profiles/r01_notes.md).
__device__ __forceinline__ float safe_rcp(float x) { return copysignf(__fdividef(1.0f, fmaxf(fabsf(x), 1e-30f)), x); }

// TransformedHittable::WorldToModel (Transform.cpp:13-20) for one chain level.
template <class M> __device__ __forceinline__ RaySpace world_to_model(const float4* __restrict__ x, RaySpace r) {
  const float4 r0 = __ldg(x + 0), r1 = __ldg(x + 1), r2 = __ldg(x + 2);
  RaySpace m;
  // vec3(inv_model * vec4(o, 1)) : (m0*x + m1*y) + (m2*z + m3*1)
  m.o.x = M::add(M::add(M::mul(r0.x, r.o.x), M::mul(r0.y, r.o.y)), M::add(M::mul(r0.z, r.o.z), r0.w));
  m.o.y = M::add(M::add(M::mul(r1.x, r.o.x), M::mul(r1.y, r.o.y)), M::add(M::mul(r1.z, r.o.z), r1.w));
  m.o.z = M::add(M::add(M::mul(r2.x, r.o.x), M::mul(r2.y, r.o.y)), M::add(M::mul(r2.z, r.o.z), r2.w));
  // normalize(mat3(inv_model) * d) : (m0*x + m1*y) + m2*z
  F3 v;
  v.x = M::add(M::add(M::mul(r0.x, r.d.x), M::mul(r0.y, r.d.y)), M::mul(r0.z, r.d.z));
  v.y = M::add(M::add(M::mul(r1.x, r.d.x), M::mul(r1.y, r.d.y)), M::mul(r1.z, r.d.z));
  v.z = M::add(M::add(M::mul(r2.x, r.d.x), M::mul(r2.y, r.d.y)), M::mul(r2.z, r.d.z));
  m.d = vnormalize<M>(v);
  return m;
}
template <class M>
__device__ __forceinline__ RaySpace to_chain_space(const DeviceScene& S, uint32_t chain_first, uint32_t chain_len, RaySpace r) {
  for (uint32_t l = 0; l < chain_len; l++) r = world_to_model<M>(S.xforms + (chain_first + l) * 6, r);
  return r;
}
// TransformedHittable::Hit epilogue (Transform.cpp:85-86), innermost level first.
template <class M>
__device__ __forceinline__ void chain_to_world(const DeviceScene& S, uint32_t chain_first, uint32_t chain_len, F3& p, F3& n) {
  for (uint32_t l = chain_len; l-- > 0;) {
    const float4* x = S.xforms + (chain_first + l) * 6;
    const float4 i0 = __ldg(x + 0), i1 = __ldg(x + 1), i2 = __ldg(x + 2);
    const float4 m0 = __ldg(x + 3), m1 = __ldg(x + 4), m2 = __ldg(x + 5);
    F3 q;
    q.x = M::add(M::add(M::mul(m0.x, p.x), M::mul(m0.y, p.y)), M::add(M::mul(m0.z, p.z), m0.w));
    q.y = M::add(M::add(M::mul(m1.x, p.x), M::mul(m1.y, p.y)), M::add(M::mul(m1.z, p.z), m1.w));
    q.z = M::add(M::add(M::mul(m2.x, p.x), M::mul(m2.y, p.y)), M::add(M::mul(m2.z, p.z), m2.w));
    p = q;
    // normal_mat = mat3(transpose(inverse(model))): (N*n)_r = (inv[0][r]*nx + inv[1][r]*ny) + inv[2][r]*nz
    F3 v;
    v.x = M::add(M::add(M::mul(i0.x, n.x), M::mul(i1.x, n.y)), M::mul(i2.x, n.z));
    v.y = M::add(M::add(M::mul(i0.y, n.x), M::mul(i1.y, n.y)), M::mul(i2.y, n.z));
    v.z = M::add(M::add(M::mul(i0.z, n.x), M::mul(i1.z, n.y)), M::mul(i2.z, n.z));
    n = vnormalize<M>(v);
  }
}

// Sphere::Hit (Sphere.cpp:7-26): returns true and the root in the OPEN interval (tmin, tmax).
template <class M>
__device__ __forceinline__ bool sphere_hit(const float4 s0, const float4 s1, F3 o, F3 d, float a, float time, float tmin,
                                           float tmax, float& t_out) {
  F3 center = ray_at<M>(make_f3(s0), make_f3(s1), time);  // center_displacement.At(r.time)
  F3 oc = vsub<M>(center, o);
  float h = vdot<M>(d, oc);
  float c = M::sub(vdot<M>(oc, oc), M::mul(s0.w, s0.w));
  float disc = M::sub(M::mul(h, h), M::mul(a, c));
  if (disc < 0.0f) return false;
  float sqrtd = M::sqrt(disc);
  float root = M::div(M::sub(h, sqrtd), a);
  if (!(tmin < root && root < tmax)) {
    root = M::div(M::add(h, sqrtd), a);
    if (!(tmin < root && root < tmax)) return false;
  }
  t_out = root;
  return true;
}

// Quad::Hit (Quad.cpp:19-35): CLOSED interval [tmin, tmax], alpha/beta in closed [0,1].
// Axis-aligned quads (u.w carries 1 + the normal's axis, set by the host: every box face, every Cornell wall): normal and w
// have one non-zero component (axis C), u and v one each, so in the reference's expressions every other product is an exact
// zero and every sum with it returns the other operand unchanged.  Evaluating only the surviving terms — n_C d_C,
// n_C o_C, component C of the two cross products, one multiply by w_C — gives the same bits with a third of the operations.
template <class M, int C>
__device__ __forceinline__ bool quad_hit_axis(const float4 nd, const float4* __restrict__ q, F3 o, F3 d, float tmin, float tmax,
                                              float& t_out) {
  constexpr int C1 = (C + 1) % 3, C2 = (C + 2) % 3;
  auto comp = [](const F3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); };
  auto comp4 = [](const float4& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); };
  const float nc = comp4(nd, C);
  const float ndd = M::mul(nc, comp(d, C));
  if (fabsf(ndd) <= 9.99999993922529e-09f) return false;
  const float t = M::div(M::sub(nd.w, M::mul(nc, comp(o, C))), ndd);
  if (!(tmin <= t && t <= tmax)) return false;
  const float4 qq = __ldg(q + 1), uu = __ldg(q + 2), vv = __ldg(q + 3), ww = __ldg(q + 4);
  // planar hit point minus q, only the two in-plane components are needed
  const float ph1 = M::sub(M::add(comp(o, C1), M::mul(comp(d, C1), t)), comp4(qq, C1));
  const float ph2 = M::sub(M::add(comp(o, C2), M::mul(comp(d, C2), t)), comp4(qq, C2));
  // (ph x v)_C = ph_C1 v_C2 - v_C1 ph_C2 ;  (u x ph)_C = u_C1 ph_C2 - ph_C1 u_C2   (glm::cross, component C)
  const float wc = comp4(ww, C);
  const float alpha = M::mul(wc, M::sub(M::mul(ph1, comp4(vv, C2)), M::mul(comp4(vv, C1), ph2)));
  const float beta = M::mul(wc, M::sub(M::mul(comp4(uu, C1), ph2), M::mul(ph1, comp4(uu, C2))));
  if (!(0.0f <= alpha && alpha <= 1.0f) || !(0.0f <= beta && beta <= 1.0f)) return false;
  t_out = t;
  return true;
}

// kAxis: take the axis-aligned path where the quad allows it.  Only the flat extend kernel does (uniform loops, registers
// to spare: Cornell +9 %); in the BVH walk and in the fused finish+shade kernel the three extra inlined variants cost more
// in registers / spills than they save (book 2 -3.6 % when enabled everywhere).
template <class M, bool kAxis = false>
__device__ __forceinline__ bool quad_hit(const float4* __restrict__ q, F3 o, F3 d, float tmin, float tmax, float& t_out) {
  const float4 nd = __ldg(q + 0);
  if (kAxis) {
    const uint32_t axis_code = __float_as_uint(__ldg(q + 2).w);  // rt2_quad.pad0
    if (axis_code == 1u) return quad_hit_axis<M, 0>(nd, q, o, d, tmin, tmax, t_out);
    if (axis_code == 2u) return quad_hit_axis<M, 1>(nd, q, o, d, tmin, tmax, t_out);
    if (axis_code == 3u) return quad_hit_axis<M, 2>(nd, q, o, d, tmin, tmax, t_out);
  }
  F3 normal = make_f3(nd);
  float ndd = vdot<M>(normal, d);
  // std::fabs(n_dot_raydir) < 1e-8 with a double literal  <=>  |x| <= float(1e-8)
  if (fabsf(ndd) <= 9.99999993922529e-09f) return false;
  float t = M::div(M::sub(nd.w, vdot<M>(normal, o)), ndd);
  if (!(tmin <= t && t <= tmax)) return false;
  const float4 qq = __ldg(q + 1), uu = __ldg(q + 2), vv = __ldg(q + 3), ww = __ldg(q + 4);
  F3 p = ray_at<M>(o, d, t);
  F3 ph = vsub<M>(p, make_f3(qq));
  float alpha = vdot<M>(make_f3(ww), vcross<M>(ph, make_f3(vv)));
  float beta = vdot<M>(make_f3(ww), vcross<M>(make_f3(uu), ph));
  if (!(0.0f <= alpha && alpha <= 1.0f) || !(0.0f <= beta && beta <= 1.0f)) return false;
  t_out = t;
  return true;
}

struct Closest {
  float t;
  uint32_t prim;      // RT2 prim ref, RT2_PRIM_NONE = miss
  int32_t instance;   // flattened instance index, -1 = world space
};

// Exact ties.  BVHNode::Hit and HittableList::Hit visit the leaves in a fixed order with a shrinking tmax (BVH.cpp:50-55,
// HittableList.cpp:8-22); Quad::Hit accepts t == tmax (closed interval, Quad.cpp:27), Sphere::Hit does not (Sphere.cpp:21).
// So among coincident faces (adjacent boxes of the book-2 grid, a box standing on the floor) the quad visited LAST by the
// reference wins, and a quad always beats a sphere at the same t.  Our traversal order is different, so the order is
// carried as a rank per quad (rt2_quad.pad1, host/scene_host.cpp: position in the reference's leaf order) and consulted
// only when t ties exactly.
__device__ __forceinline__ bool quad_wins_tie(const DeviceScene& S, float t, uint32_t ref, const Closest& best) {
  if (t < best.t) return true;
  if (best.prim == RT2_PRIM_NONE || RT2_PRIM_TYPE(best.prim) != RT2_PRIM_QUAD) return true;
  const uint32_t rank_new = __float_as_uint(__ldg(S.quads + 5 * RT2_PRIM_INDEX(ref) + 3).w);
  const uint32_t rank_old = __float_as_uint(__ldg(S.quads + 5 * RT2_PRIM_INDEX(best.prim) + 3).w);
  return rank_new > rank_old;
}

// HittableList::Hit over a short list of surface primitives (HittableList.cpp:8-22): shrinking tmax, later closed-
// interval ties win.  Used for medium boundaries.
template <class M>
__device__ __forceinline__ bool list_hit(const DeviceScene& S, uint32_t first, uint32_t count, F3 o, F3 d, float a, float time,
                                         float tmin, float tmax, float& t_out) {
  bool any = false;
  for (uint32_t i = 0; i < count; i++) {
    uint32_t ref = __ldg(S.prim_refs + first + i);
    uint32_t idx = RT2_PRIM_INDEX(ref);
    float t;
    bool h;
    if (RT2_PRIM_TYPE(ref) == RT2_PRIM_SPHERE) {
      h = sphere_hit<M>(__ldg(S.spheres + 2 * idx), __ldg(S.spheres + 2 * idx + 1), o, d, a, time, tmin, tmax, t);
    } else {
      h = quad_hit<M>(S.quads + 5 * idx, o, d, tmin, tmax, t);
    }
    if (h) {
      any = true;
      tmax = t;
      t_out = t;
    }
  }
  return any;
}

// Per-lane traversal counters (profiling build of k_traverse): algorithmic work actually done by ACTIVE lanes.
struct TravCounters {
  uint32_t box_pairs{0}, spheres{0}, quads{0}, instances{0};
  uint32_t overflow{0};  // a push found the traversal stack full (always tracked)
};

// Surface traversal of a whole ray queue by persistent warps (one lane = one ray at a time).
//
// Two-level BVH: world TLAS + one BLAS per instance.  Box tests only cull and are conservative (see cull_scale), so the
// result is the exact arg-min over leaves of the raw reported t.  Three modes share the walk:
//
//   kTravInline  the TLAS holds the instances as singleton leaves; a lane that meets one switches to the instance's model
//                space and BLAS, a sentinel on the stack switches it back (scenes with many instances).
//   kTravWorld   "instance split", pass 1: the TLAS holds surfaces only and is culled with the exact bound.  When a ray is
//                finished, its segment [0, best.t * cull_scale] is tested against the world boxes of the (few, hoisted)
//                instances — at the warp's convergence point, so all finished lanes do it together — and every touched
//                instance becomes one entry {ray, instance} of the bounce's entry queue.
//   kTravUnified ONE world-space tree whose leaves are the world surfaces and, each with a conservative world-space box,
//                every primitive of every instance (DeviceScene::inst_leaves).  An instanced leaf is still tested in its
//                instance's model space with the reference's arithmetic (the model-space ray is computed once per ray and
//                instance and parked in shared memory, next to the world ray the other leaves are tested with), so the
//                leaves and their raw t are exactly those of the two-level walks — but a ray bouncing inside an instance
//                no longer walks the whole world tree first and the instance's tree second, and the closest hit found in
//                either culls the other.
//   kTravInst    pass 2: one lane = one entry.  The exact-arithmetic world-to-model transforms (Transform.cpp:13-20) run with
//                the whole warp converged, the walk starts at the BLAS root with tmax = the ray's closest world surface,
//                and a hit is merged into the ray's slot by a 64-bit atomicMin on (ordered t, entry index).
//                Consumers read the merged record through load_closest().
//
// SIMT shape ("while-while" with dynamic fetch, after Aila & Laine 2009): the warp alternates between
//   phase 1  every lane descends interior nodes until it holds a leaf (or its ray is finished),
//   phase 2  every lane that holds a leaf intersects its primitives,
//   publish  every lane whose ray is finished writes its result,
// reconverging with __syncwarp() between the phases, so the expensive exact-arithmetic primitive tests run with many
// lanes instead of whichever lanes happen to reach a leaf in the same iteration.  When fewer than `fetch_threshold` lanes
// are busy the warp pulls new work from the queue with one atomic.  `order` (optional) is the permutation in which the
// queue is consumed.  Results go to trav_out[ray] = {t bits, prim ref, instance, has-entries flag}.
// kQuant: the walk reads the 32-byte quantised node pairs (DeviceScene::qnodes, rt_qnodes.cu) instead of the float pairs;
// the boxes are supersets of the float boxes, so the leaves reached — and the closest hit — are the same.
enum { kTravInline = 0, kTravWorld = 1, kTravInst = 2, kTravUnified = 3 };
constexpr uint32_t kFetchChunk = 64;  // queue items a warp reserves with one atomic (big queues)
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
constexpr uint32_t kTravDone = 0xFFFFFFFFu;  // `cur` of a lane whose ray is finished (carries kLeafFlag: phase 1 skips it)
constexpr uint32_t kMaxHoistedInstances = 4;

// Plumbing of the instance split (device buffers owned by the renderer).
struct SplitIO {
  uint2* entries;                  // {ray index, instance index}
  uint32_t* entry_prim;            // per entry: winning primitive inside the instance (valid when it won)
  unsigned long long* inst_best;   // per ray: ordered(t) << 32 | entry index, minimum over the ray's entries
  uint32_t* entry_count;           // entries of this bounce
  uint32_t capacity;               // entries the queue can hold (DEBUG_CHECKS)
};

// Monotone map float -> uint32 (also for negative t of caller-supplied intervals) and back.
__device__ __forceinline__ uint32_t ordered_bits(float t) {
  const uint32_t b = __float_as_uint(t);
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

// The closest SURFACE of ray i after the extend kernels: the world pass's record, replaced by the best instance entry of
// the ray when that one is closer — or exactly as close and the winner of the reference's tie rule (quad_wins_tie: the
// ranks are global, so the rule does not care which space a quad lives in).
__device__ __forceinline__ uint4 load_closest(const DeviceScene& S, const uint4* __restrict__ trav, const SplitIO& io, uint32_t i) {
  uint4 tr = trav[i];
  if (tr.w != 0u) {
    const unsigned long long b = io.inst_best[i];
    if (b != ~0ull) {
      const float t = from_ordered_bits(static_cast<uint32_t>(b >> 32));
      const float tw = __uint_as_float(tr.x);
      if (t <= tw) {
        const uint32_t e = static_cast<uint32_t>(b);
        const uint32_t prim = io.entry_prim[e];
        if (t < tw || (RT2_PRIM_TYPE(prim) == RT2_PRIM_QUAD && quad_wins_tie(S, t, prim, Closest{tw, tr.y, -1}))) {
          tr.x = __float_as_uint(t);
          tr.y = prim;
          tr.z = io.entries[e].y;
        }
      }
    }
    tr.w = 0u;
  }
  return tr;
}

template <class M, bool kCount, int kMode, bool kQuant = false>
__device__ __forceinline__ void traverse_queue(const DeviceScene& S, uint32_t n, const float4* __restrict__ ray_o,
                                               const float4* __restrict__ ray_d, float tmin, float tmax,
                                               uint32_t* __restrict__ next_ray, const uint32_t* __restrict__ order,
                                               uint4* __restrict__ trav_out, const SplitIO& io, TravCounters& cnt,
                                               int max_steps, int fetch_threshold, float* __restrict__ ms_cache = nullptr) {
  // ms_cache (kTravUnified): 14 x blockDim.x floats of shared memory — per thread the world ray (o, d, d.d) in slots 0..6 and the
  // model-space ray of the instance it last met in slots 7..13
  const unsigned kFull = 0xFFFFFFFFu;
  const unsigned lane = threadIdx.x & 31u;
  // The traversal stack lives in local memory (L1-resident).  Slot 0 is never read as an entry (it absorbs the reload after
  // the last pop), slot 1 holds a permanent bottom marker (kTravDone): popping it ends the ray without an emptiness test.
  // (A shared-memory stack — one bank per thread, so a warp-wide push or pop is a single wavefront — measured no faster:
  // 5 789 vs 5 814 Mrays/s, profiles/r02_notes.md.)
  uint32_t stack[kStackSize + 2];
  auto st_store = [&](int i, uint32_t v) { stack[i] = v; };
  auto st_load = [&](int i) -> uint32_t { return stack[i]; };
  int sp = 2;  // next free slot
  bool active = false;
  bool exhausted = false;
  uint32_t ray_idx = 0;
  uint32_t entry_idx = 0;  // kTravInst
  uint32_t chunk_pos = 0, chunk_end = 0;  // warp-uniform: the part of the queue this warp has reserved
  const bool big_queue = n >= kFetchChunk * 16u * (gridDim.x * (blockDim.x >> 5));  // >= 16 chunks per warp
  uint32_t cur = 0;
  F3 o = {0, 0, 0}, d = {0, 0, 1};
  float time = 0.0f, a = 1.0f;
  F3 inv = {0, 0, 0}, oid = {0, 0, 0};
  // kQuant: PRMT selectors of the NEAR and of the FAR plane's 16-bit field per axis
  uint32_t sel_x = 0x7104u, sel_y = 0x7104u, sel_z = 0x7104u, fx = 0x7324u, fy = 0x7324u, fz = 0x7324u;
  float cull_scale = 1.0f, cur_cull = 1.0f;  // kTravInline
  int32_t cur_inst = -1;
  Closest best{tmax, RT2_PRIM_NONE, -1};

  auto set_space = [&](F3 no, F3 nd) {
    o = no;
    d = nd;
    a = vdot<M>(d, d);
    inv = {safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z)};
    if (kQuant) {
      // fold the node grid into the slab coefficients: t = (q_base + v * q_ext - o) / d = v * (q_ext / d) + (q_base - o) / d
      oid = {(S.q_base[0] - o.x) * inv.x, (S.q_base[1] - o.y) * inv.y, (S.q_base[2] - o.z) * inv.z};
      inv = {S.q_ext[0] * inv.x, S.q_ext[1] * inv.y, S.q_ext[2] * inv.z};
      // a ray enters a slab through the min plane (low half of the word) when it travels up the axis, else through the max plane
      sel_x = inv.x < 0.0f ? 0x7324u : 0x7104u, fx = inv.x < 0.0f ? 0x7104u : 0x7324u;
      sel_y = inv.y < 0.0f ? 0x7324u : 0x7104u, fy = inv.y < 0.0f ? 0x7104u : 0x7324u;
      sel_z = inv.z < 0.0f ? 0x7324u : 0x7104u, fz = inv.z < 0.0f ? 0x7104u : 0x7324u;
    } else {
      oid = {-o.x * inv.x, -o.y * inv.y, -o.z * inv.z};
    }
  };
  // `top` mirrors the stack's top entry (slot sp - 1) in a register.  Every push / pop ends with an
  // unconditional reload of the new top, issued a whole node step before it can be needed, so a pop never waits for a load
  // and the node step below is branch-free: with 32 rays per warp some lane pushes and some lane pops in almost every step,
  // so predicating both costs no issue slots, while the branchy form ran the push at 3 and the pop at 5 of 32 lanes
  // (13 - 20 % of the kernel's warp instructions; profiles/r02_notes.md).
  // No bounds checks here: the renderer verifies at upload time that every tree (TLAS depth + 1 + deepest BLAS for the inline
  // walk) fits kStackSize and refuses the scene otherwise (Renderer::CheckTreeDepths) — a traversal can neither drop a
  // sub-tree silently nor write outside its stack.
  uint32_t top = kTravDone;
  auto pop_plain = [&]() {
    cur = top;
    sp -= 1;
    top = st_load(sp - 1);
  };
  // Pops the next entry; the bottom marker means the ray is finished (published at the end of the round).
  auto pop = [&]() {
    pop_plain();
    if (kMode == kTravInline && cur == kStackSentinel) {
      // leave the instance: back to the world-space ray (re-read instead of holding it in registers).  Instances are
      // flattened by the host, so the entry under a sentinel is never another sentinel.
      const float4 wo = ray_o[ray_idx], wd = ray_d[ray_idx];
      set_space(make_f3(wo), make_f3(wd));
      cur_inst = -1;
      cur_cull = cull_scale;
      pop_plain();
    }
  };

  while (true) {
    // ---- fetch: idle lanes take the next items of the queue ----
    // The warp owns a chunk [chunk_pos, chunk_end) of the queue and refills from it; a new chunk costs one atomic (an L2
    // round trip of ~600 cycles with every busy lane of the warp waiting) and is followed by an L2 prefetch of the chunk's
    // rays, so later refills find them on chip.  Big queues use 64-item chunks; small ones (late bounces) take exactly what
    // they need, so that no warp sits on unprocessed rays while the others have run dry.
    const unsigned idle = __ballot_sync(kFull, !active);
    if (idle) {
      if (!exhausted) {
        const uint32_t need = __popc(idle);
        const uint32_t avail = chunk_end - chunk_pos;
        uint32_t new_base = 0;
        if (avail < need) {
          const uint32_t take = big_queue ? kFetchChunk : need - avail;
          if (lane == 0) new_base = atomicAdd(next_ray, take);
          new_base = __shfl_sync(kFull, new_base, 0);
          if (big_queue && kMode != kTravInst && order == nullptr) {
            // 64 rays x (16 B origin + 16 B direction) = 8 + 8 lines of 128 B
            const uint32_t line = lane & 7u;
            const float4* src = (lane & 8u) ? ray_d : ray_o;
            if (lane < 16u && new_base + line * 8u < n) prefetch_l2(src + new_base + line * 8u);
          }
        }
        const uint32_t rank = __popc(idle & ((1u << lane) - 1u));
        const uint32_t pos = rank < avail ? chunk_pos + rank : new_base + (rank - avail);
        if (avail < need) {
          chunk_pos = new_base + (need - avail);
          chunk_end = new_base + (big_queue ? kFetchChunk : need - avail);
        } else {
          chunk_pos += need;
        }
        if (!active) {
          if (pos < n) {
            if (kMode == kTravInst) {
              // one entry: the ray in the instance's model space, bounded by its closest world-space surface
              entry_idx = pos;
              const uint2 ent = io.entries[pos];
              RT2_CHECK(ent.y < S.n_instances, kChkInstance);
              ray_idx = ent.x;
              const float4 wo = ray_o[ray_idx], wd = ray_d[ray_idx];
              time = wo.w;
              const uint4 in = __ldg(S.instances + ent.y);
              const RaySpace ms = to_chain_space<M>(S, in.x, in.y, RaySpace{make_f3(wo), make_f3(wd)});
              set_space(ms.o, ms.d);
              cur_inst = static_cast<int32_t>(ent.y);
              best.t = __uint_as_float(trav_out[ray_idx].x);
              cur = in.z;
            } else {
              const uint32_t idx = order ? order[pos] : pos;  // rt_sort.cuh: coherence order of the queue
              ray_idx = idx;
              const float4 wo = ray_o[idx], wd = ray_d[idx];
              time = wo.w;
              set_space(make_f3(wo), make_f3(wd));
              if (kMode == kTravUnified) {
                float* ws = ms_cache + threadIdx.x;
                ws[0 * blockDim.x] = o.x, ws[1 * blockDim.x] = o.y, ws[2 * blockDim.x] = o.z;
                ws[3 * blockDim.x] = d.x, ws[4 * blockDim.x] = d.y, ws[5 * blockDim.x] = d.z;
                ws[6 * blockDim.x] = a;
              }
              if (kMode == kTravInline || kMode == kTravUnified) {
                // An instanced leaf reports t in model units (= world t * |M^-1 d|): a world-space box at parameter t_w can
                // hold an instanced hit with raw t as small as t_w * sigma_min * |d|, so scale the TLAS culling bound.
                cull_scale = (S.n_instances > 0) ? fmaxf(1.0f, 1.0f / (S.min_inv_scale * sqrtf(a))) : 1.0f;
                cur_cull = cull_scale;
                cur_inst = -1;
              }
              best.t = tmax;
              cur = (kMode == kTravWorld) ? S.tlas_world_root : (kMode == kTravUnified ? S.tlas_unified_root : S.tlas_root);
            }
            best.prim = RT2_PRIM_NONE;
            best.instance = -1;
            st_store(1, kTravDone);
            sp = 2;
            top = kTravDone;
            active = true;
          }
        }
        exhausted = chunk_pos >= n;  // the warp's chunk starts beyond the queue: every later chunk does too
      }
      if (__ballot_sync(kFull, active) == 0u) break;
    }

    // ---- traverse until too few lanes are busy ----
    while (true) {
      // phase 1: interior nodes (at most max_steps per round, so that lanes holding a leaf do not wait for a long descent)
      // (warp-uniform round control — every lane stays in the loop, one ballot per step ends the round when too few lanes still
      //  walk — costs 5 % before any threshold can pay: 6 224 vs 6 565 Mrays/s; the per-lane loop exit below stays)
      for (int step = 0; step < max_steps && active && !(cur & kLeafFlag); step++) {
        RT2_CHECK(cur < S.n_node_pairs, kChkNode);
        RT2_CHECK(sp >= 2 && sp <= kStackSize, kChkStack);
        // entry / exit parameters of the two child boxes + their traversal entries
        uint32_t e0, e1;
        float near0, far0, near1, far1;
        if (kQuant) {
          // 32-byte pair, per child {minx | maxx << 16, miny | maxy << 16, minz | maxz << 16, entry}; every 16-bit field has its top
          // bit set, so PRMT {00, lo, hi, 3F} IS the float 1 + q / 32768 (rt_qnodes.cu).  The per-ray selector picks the plane the
          // ray ENTERS through (near) or leaves through (far) directly: t_near = min(t_min_plane, t_max_plane) needs no min / max.
          const uint4* np = S.qnodes + static_cast<size_t>(cur) * 2;
          const uint4 qa = __ldg(np + 0), qb = __ldg(np + 1);
          // (prmt through inline PTX: __byte_perm() masks a run-time selector with 0x7777 first — one LOP3 per use)
          auto dec = [](uint32_t w, uint32_t sel) {
            uint32_t r;
            asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(w), "r"(0x3F000000u), "r"(sel));
            return __uint_as_float(r);
          };
          if (kCount) cnt.box_pairs++;
          near0 = fmaxf(fmaxf(fmaf(dec(qa.x, sel_x), inv.x, oid.x), fmaf(dec(qa.y, sel_y), inv.y, oid.y)),
                        fmaxf(fmaf(dec(qa.z, sel_z), inv.z, oid.z), 0.0f));
          far0 = fminf(fminf(fmaf(dec(qa.x, fx), inv.x, oid.x), fmaf(dec(qa.y, fy), inv.y, oid.y)), fmaf(dec(qa.z, fz), inv.z, oid.z));
          near1 = fmaxf(fmaxf(fmaf(dec(qb.x, sel_x), inv.x, oid.x), fmaf(dec(qb.y, sel_y), inv.y, oid.y)),
                        fmaxf(fmaf(dec(qb.z, sel_z), inv.z, oid.z), 0.0f));
          far1 = fminf(fminf(fmaf(dec(qb.x, fx), inv.x, oid.x), fmaf(dec(qb.y, fy), inv.y, oid.y)), fmaf(dec(qb.z, fz), inv.z, oid.z));
          e0 = qa.w, e1 = qb.w;
        } else {
          // 64-byte pair: {min, max} corners as floats
          const float4* np = S.nodes + static_cast<size_t>(cur) * 4;
          // (one 256-bit load per node — LDG.E.256 on sm_100 — measured 4 % SLOWER than these four 128-bit loads; r02 notes)
          const float4 a0 = __ldg(np + 0), a1 = __ldg(np + 1), b0 = __ldg(np + 2), b1 = __ldg(np + 3);
          // device node format (rt_kernels.cu UploadScene / rt_lbvh.cu): .w of the min corner is the traversal entry itself
          // (interior -> child pair index; leaf -> see make_leaf_entry)
          e0 = __float_as_uint(a0.w), e1 = __float_as_uint(b0.w);
          if (kCount) cnt.box_pairs++;
          float t0x = fmaf(a0.x, inv.x, oid.x), t1x = fmaf(a1.x, inv.x, oid.x);
          float t0y = fmaf(a0.y, inv.y, oid.y), t1y = fmaf(a1.y, inv.y, oid.y);
          float t0z = fmaf(a0.z, inv.z, oid.z), t1z = fmaf(a1.z, inv.z, oid.z);
          near0 = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
          far0 = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
          t0x = fmaf(b0.x, inv.x, oid.x), t1x = fmaf(b1.x, inv.x, oid.x);
          t0y = fmaf(b0.y, inv.y, oid.y), t1y = fmaf(b1.y, inv.y, oid.y);
          t0z = fmaf(b0.z, inv.z, oid.z), t1z = fmaf(b1.z, inv.z, oid.z);
          near1 = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
          far1 = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
        }
        // conservative culling: near is shrunk by 1e-6 relative before it is compared with far and with the (scaled)
        // closest hit so far; an empty slot has NaN bounds -> far is NaN -> never entered
        const float bound = (kMode == kTravInline) ? best.t * cur_cull : (kMode == kTravUnified ? best.t * cull_scale : best.t);
        const float n0 = near0 * 0.999999f;
        const bool h0 = (n0 <= far0) && (n0 <= bound);
        const float n1 = near1 * 0.999999f;
        const bool h1 = (n1 <= far1) && (n1 <= bound);
        // branch-free step: the far child of a double hit is stored above the top (a harmless write when it is not pushed),
        // a miss takes the register copy of the top
        const bool both = h0 && h1, any = h0 || h1;
        const bool swap = both && (near1 < near0);
        const uint32_t near_child = (h0 && !swap) ? e0 : e1;
        const uint32_t far_child = swap ? e0 : e1;
        st_store(sp, far_child);
        // (prefetching the far child's node pair into L1 here, for the later pop: measured 5 931 vs 6 030 Mrays/s — dropped;
        //  predicating the store on `both` and the reload on a pop, so that only those lanes touch L1: 6 086 vs 6 182 — dropped)
        if (kMode == kTravInline) {
          if (any) {
            cur = near_child;
            sp += both ? 1 : 0;
            top = st_load(sp - 1);
          } else {
            pop();  // may have to leave an instance
          }
        } else {
          cur = any ? near_child : top;
          sp += any ? (both ? 1 : 0) : -1;
          top = st_load(sp - 1);
        }
      }
      __syncwarp();
      // phase 2: leaves
      // (Aila & Laine's postponed leaf — park a leaf met as the near child and keep walking, test it in the warp's next leaf
      //  phase — measured 6 087 vs 6 184 Mrays/s on book 2: 5 % more node visits for culling against a stale closest hit)
      const uint32_t leaf = cur;
      if (active && (leaf & kLeafFlag) && leaf != kTravDone) {
        const bool direct = (leaf & kLeafDirect) != 0u;
        const uint32_t first = leaf & 0x03FFFFFFu;
        const uint32_t count = direct ? 1u : (((leaf >> 26) & 0xFu) + 1u);
        bool entered = false;
        for (uint32_t i = 0; i < count; i++) {
          RT2_CHECK(direct || first + i < S.n_prim_refs, kChkPrimRef);
          const uint32_t ref = direct ? (leaf & 0x3FFFFFFFu) : __ldg(S.prim_refs + first + i);
          const uint32_t type = RT2_PRIM_TYPE(ref), idx = RT2_PRIM_INDEX(ref);
          RT2_CHECK(type != RT2_PRIM_SPHERE || idx < S.n_spheres, kChkSphere);
          RT2_CHECK(type != RT2_PRIM_QUAD || idx < S.n_quads, kChkQuad);
          RT2_CHECK(type != RT2_PRIM_INSTANCE || kMode != kTravUnified || idx < S.n_inst_leaves, kChkInstLeaf);
          RT2_CHECK(type != RT2_PRIM_INSTANCE || kMode != kTravInline || idx < S.n_instances, kChkInstance);
          RT2_CHECK(type <= RT2_PRIM_INSTANCE && (type != RT2_PRIM_INSTANCE || kMode == kTravUnified || kMode == kTravInline), kChkPrimRef);
          // The primitive and the ray it is tested with: a world leaf takes the world ray, an instanced leaf (kTravUnified) its
          // instance's model-space ray — then ONE sphere / quad test serves both, so the lanes of a warp split by primitive
          // type only, not by (type, space).
          uint32_t pref = ref;
          int32_t pinst = (kMode == kTravUnified) ? -1 : cur_inst;  // kTravUnified: cur_inst is the cached instance
          // kTravUnified: the world ray is parked in shared memory too (slots 0..6; the model-space ray in 7..13): it is needed by
          // leaf tests only, and the node phase gets its seven registers
          F3 ro, rd;
          float ra;
          if (kMode == kTravUnified) {
            const float* ws = ms_cache + threadIdx.x;
            ro = {ws[0 * blockDim.x], ws[1 * blockDim.x], ws[2 * blockDim.x]};
            rd = {ws[3 * blockDim.x], ws[4 * blockDim.x], ws[5 * blockDim.x]};
            ra = ws[6 * blockDim.x];
          } else {
            ro = o, rd = d, ra = a;
          }
          if (type == RT2_PRIM_INSTANCE) {
            if (kMode == kTravUnified) {
              // instanced leaf: primitive il.x of instance il.y, tested in the instance's model space (Transform.cpp:13-20,75-88)
              const uint2 il = __ldg(S.inst_leaves + idx);
              RT2_CHECK(il.y < S.n_instances, kChkInstance);
              RT2_CHECK(RT2_PRIM_TYPE(il.x) == RT2_PRIM_SPHERE ? RT2_PRIM_INDEX(il.x) < S.n_spheres : RT2_PRIM_INDEX(il.x) < S.n_quads, kChkInstLeaf);
              float* ms = ms_cache + threadIdx.x + 7 * blockDim.x;
              if (cur_inst != static_cast<int32_t>(il.y)) {
                if (kCount) cnt.instances++;
                const uint4 in = __ldg(S.instances + il.y);
                const RaySpace r = to_chain_space<M>(S, in.x, in.y, RaySpace{ro, rd});
                ms[0 * blockDim.x] = r.o.x, ms[1 * blockDim.x] = r.o.y, ms[2 * blockDim.x] = r.o.z;
                ms[3 * blockDim.x] = r.d.x, ms[4 * blockDim.x] = r.d.y, ms[5 * blockDim.x] = r.d.z;
                ms[6 * blockDim.x] = vdot<M>(r.d, r.d);
                cur_inst = static_cast<int32_t>(il.y);
              }
              ro = {ms[0 * blockDim.x], ms[1 * blockDim.x], ms[2 * blockDim.x]};
              rd = {ms[3 * blockDim.x], ms[4 * blockDim.x], ms[5 * blockDim.x]};
              ra = ms[6 * blockDim.x];
              pref = il.x;
              pinst = static_cast<int32_t>(il.y);
            } else if (kMode == kTravInline) {
              // instance leaf (always a singleton leaf of the TLAS, host/bvh_build.cpp): enter its BLAS in model space
              if (kCount) cnt.instances++;
              const uint4 in = __ldg(S.instances + idx);
              const float4 wo = ray_o[ray_idx], wd = ray_d[ray_idx];
              RaySpace ms = to_chain_space<M>(S, in.x, in.y, RaySpace{make_f3(wo), make_f3(wd)});
              set_space(ms.o, ms.d);
              cur_inst = static_cast<int32_t>(idx);
              cur_cull = 1.0f;
              st_store(sp++, kStackSentinel);
              top = kStackSentinel;
              cur = in.z;
              entered = true;
              break;
            }
          }
          const uint32_t pidx = RT2_PRIM_INDEX(pref);
          float t;
          if (RT2_PRIM_TYPE(pref) == RT2_PRIM_SPHERE) {
            if (kCount) cnt.spheres++;
            if (sphere_hit<M>(__ldg(S.spheres + 2 * pidx), __ldg(S.spheres + 2 * pidx + 1), ro, rd, ra, time, tmin, best.t, t)) {
              best.t = t;
              best.prim = pref;
              best.instance = pinst;
            }
          } else {
            if (kCount) cnt.quads++;
            if (quad_hit<M>(S.quads + 5 * pidx, ro, rd, tmin, best.t, t) && quad_wins_tie(S, t, pref, best)) {
              best.t = t;
              best.prim = pref;
              best.instance = pinst;
            }
          }
        }
        if (!entered) pop();
      }
      __syncwarp();
      // publish: finished rays, with the whole warp converged
      const bool fin = active && cur == kTravDone;
      if (kMode == kTravWorld) {
        bool has_entry = false;
        if (S.n_hoisted > 0u && __ballot_sync(kFull, fin) != 0u) {
          // which hoisted instances can still beat the closest world-space surface?  An instanced leaf reports t in model
          // units (= world t * |M^-1 d|), so the segment tested against the world box ends at best.t * cull_scale.
          const float bound = best.t * fmaxf(1.0f, 1.0f / (S.min_inv_scale * sqrtf(a)));
          for (uint32_t j = 0; j < S.n_hoisted; j++) {
            const float4 bmn = __ldg(S.inst_bounds + 2 * j), bmx = __ldg(S.inst_bounds + 2 * j + 1);
            const float t0x = fmaf(bmn.x, inv.x, oid.x), t1x = fmaf(bmx.x, inv.x, oid.x);
            const float t0y = fmaf(bmn.y, inv.y, oid.y), t1y = fmaf(bmx.y, inv.y, oid.y);
            const float t0z = fmaf(bmn.z, inv.z, oid.z), t1z = fmaf(bmx.z, inv.z, oid.z);
            const float nr = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f)) * 0.999999f;
            const float fr = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
            const bool touch = fin && (nr <= fr) && (nr <= bound);
            const unsigned tm = __ballot_sync(kFull, touch);
            if (tm) {
              const int leader = __ffs(tm) - 1;
              uint32_t base = 0;
              if (static_cast<int>(lane) == leader) base = atomicAdd(io.entry_count, __popc(tm));
              base = __shfl_sync(kFull, base, leader);
              if (touch) {
                RT2_CHECK(base + __popc(tm & ((1u << lane) - 1u)) < io.capacity, kChkEntry);
                io.entries[base + __popc(tm & ((1u << lane) - 1u))] = make_uint2(ray_idx, j);
                has_entry = true;
                if (kCount) cnt.instances++;
              }
            }
          }
        }
        if (fin) {
          trav_out[ray_idx] = make_uint4(__float_as_uint(best.t), best.prim, 0xFFFFFFFFu, has_entry ? 1u : 0u);
          if (has_entry) io.inst_best[ray_idx] = ~0ull;
          active = false;
        }
      } else if (kMode == kTravInst) {
        if (fin) {
          if (best.prim != RT2_PRIM_NONE) {
            atomicMin(io.inst_best + ray_idx, (static_cast<unsigned long long>(ordered_bits(best.t)) << 32) | entry_idx);
            io.entry_prim[entry_idx] = best.prim;
          }
          active = false;
        }
      } else {
        if (fin) {
          trav_out[ray_idx] = make_uint4(__float_as_uint(best.t), best.prim, static_cast<uint32_t>(best.instance), 0u);
          active = false;
        }
      }
      const unsigned busy = __ballot_sync(kFull, active);
      if (busy == 0u) break;
      if (!exhausted && __popc(busy) < fetch_threshold) break;
    }
  }
}

// Closest surface for TINY scenes (a Cornell box: 18 quads, 2 instances): no tree at all.  Every ray tests every world
// primitive, then enters every instance whose world box it touches and tests all of that instance's primitives — the
// same arg-min over leaves as traverse_queue, but with a loop whose trip counts and primitive types are
// identical for all 32 lanes of a warp.  On SIMT hardware that beats a BVH walk whose lanes diverge at every step
// (measured on B200: Cornell box 2.5x faster extend stage, profiles/r01_notes.md); the host picks this path when the
// scene has at most kFlatMaxPrims leaf primitives in total.
constexpr uint32_t kFlatMaxPrims = 40;

template <class M>
__device__ __forceinline__ void flat_test_range(const DeviceScene& S, uint32_t first, uint32_t last, F3 o, F3 d, float a, float time,
                                                float tmin, int32_t inst, bool lane_on, Closest& best) {
  for (uint32_t k = first; k < last; k++) {
    const uint32_t ref = __ldg(S.flat_refs + k);  // warp-uniform
    const uint32_t idx = RT2_PRIM_INDEX(ref);
    float t;
    bool h;
    if (RT2_PRIM_TYPE(ref) == RT2_PRIM_SPHERE) {
      h = sphere_hit<M>(__ldg(S.spheres + 2 * idx), __ldg(S.spheres + 2 * idx + 1), o, d, a, time, tmin, best.t, t);
    } else {
      h = quad_hit<M, true>(S.quads + 5 * idx, o, d, tmin, best.t, t);
    }
    if (h && lane_on && (RT2_PRIM_TYPE(ref) == RT2_PRIM_SPHERE || quad_wins_tie(S, t, ref, best))) {
      best.t = t;
      best.prim = ref;
      best.instance = inst;
    }
  }
}

template <class M>
__device__ __forceinline__ Closest traverse_flat(const DeviceScene& S, F3 wo, F3 wd, float time, float tmin, float tmax) {
  Closest best{tmax, RT2_PRIM_NONE, -1};
  flat_test_range<M>(S, __ldg(S.flat_offsets + 0), __ldg(S.flat_offsets + 1), wo, wd, vdot<M>(wd, wd), time, tmin, -1, true, best);
  const float ix = safe_rcp(wd.x), iy = safe_rcp(wd.y), iz = safe_rcp(wd.z);
  for (uint32_t j = 0; j < S.n_instances; j++) {
    // conservative world-box test of the whole ray (no bound by best.t: an instanced leaf reports t in model units)
    const float4 bmn = __ldg(S.inst_bounds + 2 * j), bmx = __ldg(S.inst_bounds + 2 * j + 1);
    const float ax = (bmn.x - wo.x) * ix, bx = (bmx.x - wo.x) * ix;
    const float ay = (bmn.y - wo.y) * iy, by = (bmx.y - wo.y) * iy;
    const float az = (bmn.z - wo.z) * iz, bz = (bmx.z - wo.z) * iz;
    const float tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.0f));
    const float tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    const bool touch = tn * 0.999999f <= tf;
    if (!__any_sync(__activemask(), touch)) continue;
    const uint4 in = __ldg(S.instances + j);
    const RaySpace ms = to_chain_space<M>(S, in.x, in.y, RaySpace{wo, wd});
    flat_test_range<M>(S, __ldg(S.flat_offsets + 1 + j), __ldg(S.flat_offsets + 2 + j), ms.o, ms.d, vdot<M>(ms.d, ms.d), time, tmin,
                       static_cast<int32_t>(j), touch, best);
  }
  return best;
}

// ConstantMedium::Hit (ConstantMedium.cpp:14-58), split in two: the boundary queries (deterministic, independent of the
// caller's interval) and the free-path draw against the current [tmin, tmax].  A medium in a span-1 leaf of the
// reference BVH has Hit() called twice (quirk Q2): both calls see identical boundary records, so they are computed once.
// Roots of one boundary primitive along the ray, independent of any interval: Sphere::Hit's two roots (Sphere.cpp:13-24)
// or Quad::Hit's plane parameter when the point lies inside the quad (Quad.cpp:19-35); NaN = no root.
template <class M, bool kAxis = false>
__device__ __forceinline__ void boundary_roots(const DeviceScene& S, uint32_t ref, F3 o, F3 d, float a, float time, float& r_lo,
                                               float& r_hi, bool& is_sphere) {
  const float nanv = __int_as_float(0x7FC00000);
  const uint32_t idx = RT2_PRIM_INDEX(ref);
  r_lo = nanv;
  r_hi = nanv;
  is_sphere = RT2_PRIM_TYPE(ref) == RT2_PRIM_SPHERE;
  if (is_sphere) {
    const float4 s0 = __ldg(S.spheres + 2 * idx), s1 = __ldg(S.spheres + 2 * idx + 1);
    F3 center = ray_at<M>(make_f3(s0), make_f3(s1), time);
    F3 oc = vsub<M>(center, o);
    float h = vdot<M>(d, oc);
    float c = M::sub(vdot<M>(oc, oc), M::mul(s0.w, s0.w));
    float disc = M::sub(M::mul(h, h), M::mul(a, c));
    if (disc < 0.0f) return;
    float sqrtd = M::sqrt(disc);
    r_lo = M::div(M::sub(h, sqrtd), a);
    r_hi = M::div(M::add(h, sqrtd), a);
  } else {
    float t;
    if (quad_hit<M, kAxis>(S.quads + 5 * idx, o, d, -kFltMax, kFltMax, t)) r_lo = t;
  }
}
// What <primitive>::Hit(r, Interval(lo, kInfinity)) reports for the roots above: spheres use the open interval and prefer
// the near root (Sphere.cpp:19-24), quads the closed one (Quad.cpp:27).  NaN = miss.
__device__ __forceinline__ float boundary_candidate(float r_lo, float r_hi, bool is_sphere, float lo) {
  const float nanv = __int_as_float(0x7FC00000);
  if (is_sphere) {
    if (lo < r_lo && r_lo < kFltMax) return r_lo;
    if (lo < r_hi && r_hi < kFltMax) return r_hi;
    return nanv;
  }
  return (lo <= r_lo && r_lo <= kFltMax) ? r_lo : nanv;
}

// The two boundary queries of ConstantMedium::Hit (ConstantMedium.cpp:18-24):
//   boundary_->Hit(r, Interval::kUniverse, rec1);  boundary_->Hit(r, Interval(rec1.t + 0.0001, kInfinity), rec2)
// A closest-hit query over a list returns the minimum of the per-primitive candidates (HittableList.cpp:8-22: shrinking
// tmax), and a primitive's roots do not depend on the interval, so for short boundaries (a sphere, a box of 6 quads) the
// roots are computed ONCE and both queries are answered from them — same values bit for bit, half the arithmetic.
constexpr uint32_t kBoundaryRootsMax = 6;
template <class M, bool kAxis = false>
__device__ __forceinline__ bool medium_boundary(const DeviceScene& S, const uint4 m0, F3 o, F3 d, float a, float time, float& t1,
                                                float& t2) {
  if (m0.w == 1u) {  // a single primitive (the usual sphere-bounded medium)
    float lo, hi;
    bool sph;
    boundary_roots<M, kAxis>(S, __ldg(S.prim_refs + m0.z), o, d, a, time, lo, hi, sph);
    t1 = boundary_candidate(lo, hi, sph, -kFltMax);
    if (!(t1 <= kFltMax)) return false;
    const float t1_eps = static_cast<float>(static_cast<double>(t1) + 0.0001);
    t2 = boundary_candidate(lo, hi, sph, t1_eps);
    return t2 <= kFltMax;
  }
  if (m0.w <= kBoundaryRootsMax) {
    float lo[kBoundaryRootsMax], hi[kBoundaryRootsMax];
    bool sph[kBoundaryRootsMax];
    float q1 = kFltMax;
    bool any1 = false;
#pragma unroll
    for (uint32_t k = 0; k < kBoundaryRootsMax; k++) {
      lo[k] = hi[k] = __int_as_float(0x7FC00000);
      sph[k] = false;
      if (k < m0.w) {
        boundary_roots<M, kAxis>(S, __ldg(S.prim_refs + m0.z + k), o, d, a, time, lo[k], hi[k], sph[k]);
        const float c = boundary_candidate(lo[k], hi[k], sph[k], -kFltMax);
        if (c <= q1) {  // NaN compares false
          q1 = c;
          any1 = true;
        }
      }
    }
    if (!any1) return false;
    t1 = q1;
    // the sum is formed in double, then stored as float
    const float t1_eps = static_cast<float>(static_cast<double>(t1) + 0.0001);
    float q2 = kFltMax;
    bool any2 = false;
#pragma unroll
    for (uint32_t k = 0; k < kBoundaryRootsMax; k++) {
      const float c = boundary_candidate(lo[k], hi[k], sph[k], t1_eps);
      if (c <= q2) {
        q2 = c;
        any2 = true;
      }
    }
    t2 = q2;
    return any2;
  }
  // long boundary lists: two sequential queries
  if (!list_hit<M>(S, m0.z, m0.w, o, d, a, time, -kFltMax, kFltMax, t1)) return false;
  const float t1_eps = static_cast<float>(static_cast<double>(t1) + 0.0001);
  return list_hit<M>(S, m0.z, m0.w, o, d, a, time, t1_eps, kFltMax, t2);
}
template <class M>
__device__ __forceinline__ bool medium_draw(float hit_dist, float ray_len, float t1, float t2, float tmin, float tmax, float& t_out) {
  t1 = fmaxf(t1, tmin);
  t2 = fminf(t2, tmax);
  if (t1 >= t2) return false;
  t1 = fmaxf(t1, 0.0f);
  const float dist_inside = M::mul(M::sub(t2, t1), ray_len);
  if (hit_dist > dist_inside) return false;
  t_out = M::add(t1, M::div(hit_dist, ray_len));
  return true;
}

struct HitOut {
  F3 p;
  float t;
  float u, v;  // HitRecord::uv (only computed when the scene has image textures)
  F3 n;
  int32_t material;  // -1 = miss
  bool front_face;
  uint32_t prim;
  int32_t instance;
};

// Constant media (few per scene) against the closest surface so far: each draws its free path (ConstantMedium.cpp:14-58)
// against the current best raw t and shrinks it when it scatters first.  medium_hit = index of the winner or -1.
// kSimple: the host guarantees that every medium has a ONE-primitive boundary in world space (no instance chain) — the
// sphere-bounded media of book 2.  The chain transforms, the 6-root and the list boundary code then drop out of the kernel
// (the fused finish + shade kernel is instruction-cache bound: profiles/r02_notes.md).
template <class M, bool kAxis = false, bool kSimple = false>
__device__ __forceinline__ void media_sample(const DeviceScene& S, F3 wo, F3 wd, float time, float tmin, const RngKey& key,
                                             uint32_t bounce, Closest& best, int32_t& medium_hit) {
  medium_hit = -1;
  uint4 draw = make_uint4(0, 0, 0, 0);
  bool have_draw = false;
  {
    for (uint32_t m = 0; m < S.n_media; m++) {
      const uint4 m0 = __ldg(S.media + 2 * m), m1 = __ldg(S.media + 2 * m + 1);
      RaySpace rs{wo, wd};
      if (!kSimple) rs = to_chain_space<M>(S, m1.x, m1.y, rs);
      const float a = vdot<M>(rs.d, rs.d);
      const float ray_len = M::sqrt(a);  // glm::length(r.direction)
      // one Philox call serves two media: .xy for even m, .zw for odd m
      if ((m & 1u) == 0u || !have_draw) {
        draw = rng_draw(key, bounce, kStreamMedium + (m >> 1));
        have_draw = true;
      }
      const uint32_t xi1 = (m & 1u) ? draw.z : draw.x, xi2 = (m & 1u) ? draw.w : draw.y;
      // hit_distance = neg_inv_density * std::log(RandReal()) (ConstantMedium.cpp:42), one draw per Hit() call
      const float hd1 = M::mul(__uint_as_float(m0.x), logf(u01(xi1)));
      const float hd2 = m1.z ? M::mul(__uint_as_float(m0.x), logf(u01(xi2))) : kFltMax;
      // Early out (no boundary roots needed): the segment inside the medium is never longer than the caller's
      // interval, so a free path beyond (tmax - max(tmin, 0)) * |d| (+ rounding slack) cannot scatter.  Thin fog that
      // encloses the whole scene (book 2: r = 5000, density 1e-4) is rejected here for ~9 rays out of 10.
      const float span = (best.t - fmaxf(tmin, 0.0f)) * ray_len * 1.00001f;
      if (hd1 > span && hd2 > span) continue;
      float t1, t2;
      if (kSimple) {
        float lo, hi;
        bool sph;
        boundary_roots<M, kAxis>(S, __ldg(S.prim_refs + m0.z), rs.o, rs.d, a, time, lo, hi, sph);
        t1 = boundary_candidate(lo, hi, sph, -kFltMax);
        if (!(t1 <= kFltMax)) continue;
        t2 = boundary_candidate(lo, hi, sph, static_cast<float>(static_cast<double>(t1) + 0.0001));
        if (!(t2 <= kFltMax)) continue;
      } else if (!medium_boundary<M, kAxis>(S, m0, rs.o, rs.d, a, time, t1, t2)) {
        continue;
      }
      float t;
      if (medium_draw<M>(hd1, ray_len, t1, t2, tmin, best.t, t)) {
        best.t = t;
        medium_hit = static_cast<int32_t>(m);
      }
      if (m1.z) {  // span-1 leaf of the reference BVH: Hit() runs twice, the second against the shrunken interval
        if (medium_draw<M>(hd2, ray_len, t1, t2, tmin, best.t, t)) {
          best.t = t;
          medium_hit = static_cast<int32_t>(m);
        }
      }
    }
  }
}

// The winner's hit record: a medium scatter (medium_hit >= 0, at best.t), the closest surface best.prim, or a miss.
// kImages = false drops the (u, v) code (acosf / atan2f: ~3 KB) from kernels specialised for scenes without image textures.
template <class M, bool kSimple = false, bool kImages = true>
__device__ __forceinline__ void finish_record(const DeviceScene& S, F3 wo, F3 wd, float time, Closest best, int32_t medium_hit, HitOut& out) {
  out.t = best.t;
  out.prim = best.prim;
  out.instance = best.instance;
  out.u = out.v = 0.0f;
  if (medium_hit >= 0) {
    RT2_CHECK(static_cast<uint32_t>(medium_hit) < S.n_media, kChkMedium);
    const uint4 m0 = __ldg(S.media + 2 * medium_hit), m1 = __ldg(S.media + 2 * medium_hit + 1);
    RaySpace rs{wo, wd};
    if (!kSimple) rs = to_chain_space<M>(S, m1.x, m1.y, rs);
    F3 p = ray_at<M>(rs.o, rs.d, best.t);
    F3 n = {1.0f, 0.0f, 0.0f};  // "both arbitrary", ConstantMedium.cpp:52-53
    if (!kSimple) chain_to_world<M>(S, m1.x, m1.y, p, n);
    out.p = p;
    out.n = n;
    out.front_face = true;
    out.material = static_cast<int32_t>(m0.y);
    out.prim = (RT2_PRIM_MEDIUM << 28) | static_cast<uint32_t>(medium_hit);
    out.instance = -1;
    return;
  }
  if (best.prim == RT2_PRIM_NONE) {
    out.material = -1;
    out.p = {0, 0, 0};
    out.n = {0, 0, 0};
    out.front_face = false;
    out.t = 0.0f;
    return;
  }
  // rebuild the winner's record in its own space with the reference's arithmetic, then map to world
  RaySpace rs{wo, wd};
  uint32_t chain_first = 0, chain_len = 0;
  if (best.instance >= 0) {
    const uint4 in = __ldg(S.instances + best.instance);
    chain_first = in.x;
    chain_len = in.y;
    rs = to_chain_space<M>(S, chain_first, chain_len, rs);
  }
  const uint32_t idx = RT2_PRIM_INDEX(best.prim);
  RT2_CHECK(RT2_PRIM_TYPE(best.prim) == RT2_PRIM_SPHERE ? idx < S.n_spheres : (RT2_PRIM_TYPE(best.prim) == RT2_PRIM_QUAD && idx < S.n_quads), kChkPrimRef);
  RT2_CHECK(best.instance < static_cast<int32_t>(S.n_instances), kChkInstance);
  F3 p = ray_at<M>(rs.o, rs.d, best.t);
  F3 outward;
  uint32_t mat;
  if (RT2_PRIM_TYPE(best.prim) == RT2_PRIM_SPHERE) {
    const float4 s0 = __ldg(S.spheres + 2 * idx), s1 = __ldg(S.spheres + 2 * idx + 1);
    F3 center = ray_at<M>(make_f3(s0), make_f3(s1), time);
    outward = vdivs<M>(vsub<M>(p, center), s0.w);  // Sphere.cpp:32
    mat = __float_as_uint(s1.w);
  } else {
    const float4 nd = __ldg(S.quads + 5 * idx), qq = __ldg(S.quads + 5 * idx + 1);
    outward = make_f3(nd);
    mat = __float_as_uint(qq.w);
  }
  if (kImages && S.n_images) {
    if (RT2_PRIM_TYPE(best.prim) == RT2_PRIM_SPHERE) {
      // Sphere::GetUV(outward_normal) (Sphere.cpp:34,39-43)
      const float theta = acosf(-outward.y);
      const float phi = atan2f(-outward.z, outward.x) + 3.14159265358979323846f;
      out.u = phi / (2.0f * 3.14159265358979323846f);
      out.v = theta / 3.14159265358979323846f;
    } else {
      // Quad::IsInterior stores the planar coordinates (Quad.cpp:15): alpha = w . ((p - q) x v), beta = w . (u x (p - q))
      const float4 qq = __ldg(S.quads + 5 * idx + 1), uu = __ldg(S.quads + 5 * idx + 2), vv = __ldg(S.quads + 5 * idx + 3),
                   ww = __ldg(S.quads + 5 * idx + 4);
      const F3 ph = vsub<M>(p, make_f3(qq));
      out.u = vdot<M>(make_f3(ww), vcross<M>(ph, make_f3(vv)));
      out.v = vdot<M>(make_f3(ww), vcross<M>(make_f3(uu), ph));
    }
  }
  // HitRecord::SetFaceNormal (HitRecord.hpp:17-20)
  bool ff = vdot<M>(rs.d, outward) < 0.0f;
  F3 n = ff ? outward : vneg(outward);
  chain_to_world<M>(S, chain_first, chain_len, p, n);
  out.p = p;
  out.n = n;
  out.front_face = ff;
  out.material = static_cast<int32_t>(mat);
}

// Second half of the closest-hit query ≡ scene.hittable_list.Hit(r, Interval{tmin, tmax}, rec) (RayTracer.cpp:25): given the
// closest SURFACE (traverse_queue), sample the constant media against it and build the winner's record.
template <class M, bool kSimple = false, bool kImages = true>
__device__ __forceinline__ void finish_hit(const DeviceScene& S, F3 wo, F3 wd, float time, float tmin, Closest best,
                                           const RngKey& key, uint32_t bounce, bool skip_media, HitOut& out) {
  int32_t medium_hit = -1;
  if (!skip_media) media_sample<M, false, kSimple>(S, wo, wd, time, tmin, key, bounce, best, medium_hit);
  finish_record<M, kSimple, kImages>(S, wo, wd, time, best, medium_hit, out);
}

}  // namespace rt2dev
