// 4-wide BVH with quantised child boxes for scenes whose tree does not fit the caches (the 1 M – 10 M sphere
// stress scene).  The reference has no counterpart (its BVH is a binary pointer tree, BVH.cpp:10-55); closest-hit results do not
// depend on the tree.
//
// Why: on the 10 M-sphere scene the binary walk visits 48 node pairs of 64 B per ray and is bound by cache capacity / L2
// traffic (more resident warps make it SLOWER, profiles/r01_notes.md).  A wide node holds the four grandchildren of a node
// pair in ONE 64-byte record — a common origin and per-axis scale in float, 8-bit conservative child boxes, four links —
// so a ray fetches half as many records of the same size: half the bytes and half the dependent round trips.
//
//   record (4 x uint4 = 64 B):  v0 = {lo.x, lo.y, lo.z, scale.x}   v1 = {scale.y, scale.z, qlo.x[4], qlo.y[4]}
//                               v2 = {qlo.z[4], qhi.x[4], qhi.y[4], qhi.z[4]}        v3 = {link[4]}
//   child k's box = lo + q * scale (fmaf, the same expression at build time, so the rounding the builder checked is the rounding
//   the walk sees); an empty slot has qlo = 255 > qhi = 0.  link = traversal entry as in rt_trace.cuh (wide-node index = the
//   node-pair index it was collapsed from, or a one-primitive leaf).
#pragma once
#include "rt_trace.cuh"

namespace rt2dev {

struct WideScene {
  const uint4* __restrict__ nodes4;  // 4 x uint4 per wide node, indexed by node-pair index
  uint32_t root;
};

__device__ __forceinline__ float wide_decode(uint32_t packed, int k, float scale, float lo) {
  return fmaf(static_cast<float>((packed >> (8 * k)) & 0xFFu), scale, lo);
}

// One thread per node pair: collapse the pair's children (or, for interior children, THEIR children) into one wide record.
__global__ void __launch_bounds__(256) k_wide_collapse(const float4* __restrict__ nodes2, uint32_t n_pairs, uint4* __restrict__ nodes4) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  float smin[4][3], smax[4][3];
  uint32_t link[4];
  int n = 0;
  auto add = [&](const float4 mn, const float4 mx) {
    if (!(mn.x <= mx.x)) return;  // empty slot (NaN bounds)
    smin[n][0] = mn.x, smin[n][1] = mn.y, smin[n][2] = mn.z;
    smax[n][0] = mx.x, smax[n][1] = mx.y, smax[n][2] = mx.z;
    link[n] = __float_as_uint(mn.w);
    n++;
  };
  for (int c = 0; c < 2; c++) {
    const float4 mn = nodes2[static_cast<size_t>(p) * 4 + 2 * c], mx = nodes2[static_cast<size_t>(p) * 4 + 2 * c + 1];
    if (!(mn.x <= mx.x)) continue;
    const uint32_t e = __float_as_uint(mn.w);
    if (e & kLeafFlag) {
      add(mn, mx);
    } else {
      const float4* q = nodes2 + static_cast<size_t>(e) * 4;
      add(q[0], q[1]);
      add(q[2], q[3]);
    }
  }
  float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0}, scale[3];
  for (int a = 0; a < 3; a++) {
    if (n) lo[a] = smin[0][a], hi[a] = smax[0][a];
    for (int k = 1; k < n; k++) lo[a] = fminf(lo[a], smin[k][a]), hi[a] = fmaxf(hi[a], smax[k][a]);
    scale[a] = fmaxf((hi[a] - lo[a]) * (1.000001f / 255.0f), 1e-30f);
  }
  uint32_t qlo[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, qhi[3] = {0u, 0u, 0u};  // empty slots: 255 > 0
  for (int k = 0; k < n; k++) {
    for (int a = 0; a < 3; a++) {
      int l = static_cast<int>(floorf((smin[k][a] - lo[a]) / scale[a]));
      l = l < 0 ? 0 : (l > 255 ? 255 : l);
      while (l > 0 && fmaf(static_cast<float>(l), scale[a], lo[a]) > smin[k][a]) l--;
      int h = static_cast<int>(ceilf((smax[k][a] - lo[a]) / scale[a]));
      h = h < 0 ? 0 : (h > 255 ? 255 : h);
      while (h < 255 && fmaf(static_cast<float>(h), scale[a], lo[a]) < smax[k][a]) h++;
      qlo[a] = (qlo[a] & ~(0xFFu << (8 * k))) | (static_cast<uint32_t>(l) << (8 * k));
      qhi[a] = (qhi[a] & ~(0xFFu << (8 * k))) | (static_cast<uint32_t>(h) << (8 * k));
    }
  }
  uint4* out = nodes4 + static_cast<size_t>(p) * 4;
  out[0] = make_uint4(__float_as_uint(lo[0]), __float_as_uint(lo[1]), __float_as_uint(lo[2]), __float_as_uint(scale[0]));
  out[1] = make_uint4(__float_as_uint(scale[1]), __float_as_uint(scale[2]), qlo[0], qlo[1]);
  out[2] = make_uint4(qlo[2], qhi[0], qhi[1], qhi[2]);
  out[3] = make_uint4(n > 0 ? link[0] : 0u, n > 1 ? link[1] : 0u, n > 2 ? link[2] : 0u, n > 3 ? link[3] : 0u);
}

// Closest surface over the wide tree: same contract and SIMT shape as traverse_queue (persistent warps, dynamic fetch,
// node phase / leaf phase), world space only (the wide tree is built for scenes without instances).
template <class M, bool kCount, int kFetchThreshold>
__device__ __forceinline__ void traverse_queue_wide(const DeviceScene& S, const WideScene W, uint32_t n, const float4* __restrict__ ray_o,
                                                    const float4* __restrict__ ray_d, float tmin, float tmax,
                                                    uint32_t* __restrict__ next_ray, uint4* __restrict__ trav_out, TravCounters& cnt,
                                                    int max_steps) {
  const unsigned kFull = 0xFFFFFFFFu;
  const unsigned lane = threadIdx.x & 31u;
  constexpr int kWideStack = 96;  // up to 3 pushes per level of a tree half as deep as the binary one
  uint32_t stack[kWideStack];
  int sp = 0;
  bool active = false, exhausted = false;
  uint32_t ray_idx = 0, cur = 0;
  F3 o = {0, 0, 0}, d = {0, 0, 1}, inv = {0, 0, 0}, oid = {0, 0, 0};
  float time = 0.0f, a = 1.0f;
  Closest best{tmax, RT2_PRIM_NONE, -1};

  auto pop = [&]() {
    if (sp == 0) {
      trav_out[ray_idx] = make_uint4(__float_as_uint(best.t), best.prim, 0xFFFFFFFFu, 0u);
      active = false;
      return;
    }
    cur = stack[--sp];
  };

  while (true) {
    const unsigned idle = __ballot_sync(kFull, !active);
    if (idle) {
      if (!exhausted) {
        const int leader = __ffs(idle) - 1;
        uint32_t base = 0;
        if (static_cast<int>(lane) == leader) base = atomicAdd(next_ray, __popc(idle));
        base = __shfl_sync(kFull, base, leader);
        if (!active) {
          const uint32_t pos = base + __popc(idle & ((1u << lane) - 1u));
          if (pos < n) {
            ray_idx = pos;
            const float4 wo = ray_o[pos], wd = ray_d[pos];
            time = wo.w;
            o = make_f3(wo);
            d = make_f3(wd);
            a = vdot<M>(d, d);
            inv = {safe_rcp(d.x), safe_rcp(d.y), safe_rcp(d.z)};
            oid = {-o.x * inv.x, -o.y * inv.y, -o.z * inv.z};
            best.t = tmax;
            best.prim = RT2_PRIM_NONE;
            sp = 0;
            cur = W.root;
            active = true;
          }
        }
        exhausted = (base + __popc(idle)) >= n;
      }
      if (__ballot_sync(kFull, active) == 0u) break;
    }
    while (true) {
      // phase 1: wide interior nodes
      for (int step = 0; step < max_steps && active && !(cur & kLeafFlag); step++) {
        const uint4* np = W.nodes4 + static_cast<size_t>(cur) * 4;
        const uint4 v0 = __ldg(np + 0), v1 = __ldg(np + 1), v2 = __ldg(np + 2), v3 = __ldg(np + 3);
        if (kCount) cnt.box_pairs += 2;
        const float lox = __uint_as_float(v0.x), loy = __uint_as_float(v0.y), loz = __uint_as_float(v0.z);
        const float sx = __uint_as_float(v0.w), sy = __uint_as_float(v1.x), sz = __uint_as_float(v1.y);
        float nearv[4];
        uint32_t ent[4] = {v3.x, v3.y, v3.z, v3.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {
          const float t0x = fmaf(wide_decode(v1.z, k, sx, lox), inv.x, oid.x), t1x = fmaf(wide_decode(v2.y, k, sx, lox), inv.x, oid.x);
          const float t0y = fmaf(wide_decode(v1.w, k, sy, loy), inv.y, oid.y), t1y = fmaf(wide_decode(v2.z, k, sy, loy), inv.y, oid.y);
          const float t0z = fmaf(wide_decode(v2.x, k, sz, loz), inv.z, oid.z), t1z = fmaf(wide_decode(v2.w, k, sz, loz), inv.z, oid.z);
          const float nr = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.0f));
          const float fr = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fmaxf(t0z, t1z));
          const float nn = nr * 0.999999f;
          // an empty slot decodes to lo + 255 s .. lo: the ray would have to be inside an inverted box on every axis
          const bool empty = ((v1.z >> (8 * k)) & 0xFFu) > ((v2.y >> (8 * k)) & 0xFFu);
          nearv[k] = (!empty && nn <= fr && nn <= best.t) ? nr : kFltMax;
        }
        // sort the four (near, entry) pairs, nearest first (5 compare-exchanges); misses carry near = FLT_MAX
        auto cswap = [&](int i, int j) {
          if (nearv[j] < nearv[i]) {
            const float tf = nearv[i];
            nearv[i] = nearv[j];
            nearv[j] = tf;
            const uint32_t te = ent[i];
            ent[i] = ent[j];
            ent[j] = te;
          }
        };
        cswap(0, 1);
        cswap(2, 3);
        cswap(0, 2);
        cswap(1, 3);
        cswap(1, 2);
        if (nearv[0] == kFltMax) {
          pop();
        } else {
          cur = ent[0];
          // push the others farthest first, so that the nearest is popped first
          if (nearv[3] != kFltMax) {
            if (sp < kWideStack) stack[sp++] = ent[3];
            else cnt.overflow = 1u;
          }
          if (nearv[2] != kFltMax) {
            if (sp < kWideStack) stack[sp++] = ent[2];
            else cnt.overflow = 1u;
          }
          if (nearv[1] != kFltMax) {
            if (sp < kWideStack) stack[sp++] = ent[1];
            else cnt.overflow = 1u;
          }
        }
      }
      __syncwarp();
      // phase 2: one-primitive leaves (the device LBVH emits nothing else)
      if (active && (cur & kLeafFlag)) {
        const bool direct = (cur & kLeafDirect) != 0u;
        const uint32_t first = cur & 0x03FFFFFFu;
        const uint32_t count = direct ? 1u : (((cur >> 26) & 0xFu) + 1u);
        for (uint32_t i = 0; i < count; i++) {
          const uint32_t ref = direct ? (cur & 0x3FFFFFFFu) : __ldg(S.prim_refs + first + i);
          const uint32_t idx = RT2_PRIM_INDEX(ref);
          float t;
          bool h;
          if (RT2_PRIM_TYPE(ref) == RT2_PRIM_SPHERE) {
            if (kCount) cnt.spheres++;
            h = sphere_hit<M>(__ldg(S.spheres + 2 * idx), __ldg(S.spheres + 2 * idx + 1), o, d, a, time, tmin, best.t, t);
          } else {
            if (kCount) cnt.quads++;
            h = quad_hit<M>(S.quads + 5 * idx, o, d, tmin, best.t, t);
          }
          if (h) {
            best.t = t;
            best.prim = ref;
          }
        }
        pop();
      }
      __syncwarp();
      const unsigned busy = __ballot_sync(kFull, active);
      if (busy == 0u) break;
      if (!exhausted && __popc(busy) < kFetchThreshold) break;
    }
  }
}

}  // namespace rt2dev
