// GPU LBVH builder (north_star: "or a GPU LBVH/wide-BVH build"): builds the 32-byte-node / 64-byte-pair
// BVH2 of include/rt2.h on the device from per-primitive boxes.  It replaces the reference's host-side median-split
// recursion (src/cpu_raytrace/BVH.cpp:10-31, one std::sort per level) for scenes where a host build is too slow — the
// synthetic 1M-10M sphere stress scene (BASELINE config 5).  Closest-hit results do not depend on the tree.
//
// Pipeline (all hand-written kernels, HBM-bound integer work; no CUB / Thrust):
//   k_lbvh_bounds      centroid bounds (block reduction + ordered-int atomics)
//   k_lbvh_morton      63-bit Morton code of each centroid (21 bits per axis), value = primitive slot
//   k_radix_*          LSD radix sort of (key, value), 8-bit digits over the bits in use: per-tile histograms, one scan, stable scatter
//   k_lbvh_gather      leaf boxes / primitive references in sorted order
//   k_lbvh_hierarchy   Karras 2012: one thread per internal node finds its key range and split
//   k_lbvh_refit       bottom-up boxes, second arrival at a node does the union
//   k_lbvh_emit        internal node i -> node pair i {left child box + link, right child box + link}
// Internal node i of the Karras tree IS node pair (base + i); leaves hold exactly one primitive.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <string>

#include "rt_lbvh.hpp"

namespace rt2dev {

constexpr int kLbvhBlock = 256;
constexpr int kRadixBits = 8;
constexpr int kRadixBins = 1 << kRadixBits;
constexpr int kKeysPerThread = 8;
constexpr int kTile = kLbvhBlock * kKeysPerThread;  // 2048 keys per block

__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

// bounds[0..2] = min centroid, bounds[3..5] = max centroid (ordered-int encoding), initialised to +-FLT_MAX by the caller;
// moments[0..2] = sum of centroids, moments[3..5] = sum of squares (double), zeroed by the caller.
__global__ void __launch_bounds__(kLbvhBlock) k_lbvh_bounds(const rt2::BuildPrim* __restrict__ prims, uint32_t n, int* __restrict__ bounds,
                                                            double* __restrict__ moments) {
  float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
  double s1[3] = {0, 0, 0}, s2[3] = {0, 0, 0};
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const rt2::BuildPrim p = prims[i];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float c = 0.5f * (p.bmin[k] + p.bmax[k]);
      mn[k] = fminf(mn[k], c);
      mx[k] = fmaxf(mx[k], c);
      s1[k] += c;
      s2[k] += static_cast<double>(c) * c;
    }
  }
#pragma unroll
  for (int k = 0; k < 3; k++) {
    for (int off = 16; off > 0; off >>= 1) {
      mn[k] = fminf(mn[k], __shfl_down_sync(0xFFFFFFFFu, mn[k], off));
      mx[k] = fmaxf(mx[k], __shfl_down_sync(0xFFFFFFFFu, mx[k], off));
      s1[k] += __shfl_down_sync(0xFFFFFFFFu, s1[k], off);
      s2[k] += __shfl_down_sync(0xFFFFFFFFu, s2[k], off);
    }
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      atomicMin(&bounds[k], float_to_ordered(mn[k]));
      atomicMax(&bounds[3 + k], float_to_ordered(mx[k]));
      atomicAdd(&moments[k], s1[k]);
      atomicAdd(&moments[3 + k], s2[k]);
    }
  }
}

// Morton grid = mean +- 3 sigma of the centroids, clamped to their true bounds: a single far-away giant (the r = 1e5 ground
// sphere of the stress scene) would otherwise squeeze every other primitive into a handful of cells along one axis.
// Every axis gets its own cell size (2^21 cells over its own extent).  An ISOTROPIC grid (one cell size for all axes, so that
// a flat scene is not cut as often along its thin axis) was measured WORSE on the 2000 x 200 x 2000 stress scene: 35.6 vs
// 30.0 node pairs per ray at 1 M spheres, 49.9 vs 48.0 at 10 M (profiles/r02_notes.md) — the early cuts across the thin
// axis are what separates the rays that skim over the slab from the spheres inside it.
// grid[0..2] = origin, grid[3..5] = cells per unit.
constexpr int kMortonBitsPerAxis = 21;
__global__ void k_lbvh_grid(uint32_t n, const int* __restrict__ bounds, const double* __restrict__ moments, float* __restrict__ grid) {
  if (threadIdx.x >= 3 || blockIdx.x != 0) return;
  const int k = threadIdx.x;
  const double mean = moments[k] / n;
  const double var = fmax(moments[3 + k] / n - mean * mean, 0.0);
  const double sd = sqrt(var);
  const float lo = fmaxf(ordered_to_float(bounds[k]), static_cast<float>(mean - 3.0 * sd));
  const float hi = fminf(ordered_to_float(bounds[3 + k]), static_cast<float>(mean + 3.0 * sd));
  grid[k] = lo;
  grid[3 + k] = hi > lo ? static_cast<float>((1u << kMortonBitsPerAxis) - 1u) / (hi - lo) : 0.0f;
}

// spreads the low 21 bits of v to every third bit
__device__ __forceinline__ unsigned long long expand_bits21(uint32_t v) {
  unsigned long long x = v & 0x1FFFFFull;
  x = (x | (x << 32)) & 0x001F00000000FFFFull;
  x = (x | (x << 16)) & 0x001F0000FF0000FFull;
  x = (x | (x << 8)) & 0x100F00F00F00F00Full;
  x = (x | (x << 4)) & 0x10C30C30C30C30C3ull;
  x = (x | (x << 2)) & 0x1249249249249249ull;
  return x;
}

__global__ void __launch_bounds__(kLbvhBlock) k_lbvh_morton(const rt2::BuildPrim* __restrict__ prims, uint32_t n, const float* __restrict__ grid,
                                                            unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals,
                                                            unsigned long long* __restrict__ key_or) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long key = 0;
  if (i < n) {
    const rt2::BuildPrim p = prims[i];
    uint32_t q[3];
    const float top = static_cast<float>((1u << kMortonBitsPerAxis) - 1u);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const float c = 0.5f * (p.bmin[k] + p.bmax[k]);
      const float u = fminf(fmaxf((c - grid[k]) * grid[3 + k], 0.0f), top);  // outliers clamp to the border cells
      q[k] = static_cast<uint32_t>(u);
    }
    key = (expand_bits21(q[0]) << 2) | (expand_bits21(q[1]) << 1) | expand_bits21(q[2]);
    keys[i] = key;
    vals[i] = i;
  }
  // OR of all keys: radix passes over digits that are zero in every key are skipped by the host
  for (int off = 16; off > 0; off >>= 1) key |= __shfl_down_sync(0xFFFFFFFFu, key, off);
  if ((threadIdx.x & 31) == 0 && key) atomicOr(key_or, key);
}

// ---- radix sort ---------------------------------------------------------------------------------------------------
// hist layout: hist[bin * n_tiles + tile] so that one exclusive scan over the whole array yields global offsets.
__global__ void __launch_bounds__(kLbvhBlock) k_radix_hist(const unsigned long long* __restrict__ keys, uint32_t n, int shift, uint32_t n_tiles,
                                                           uint32_t* __restrict__ hist) {
  __shared__ uint32_t sh[kRadixBins];
  for (int b = threadIdx.x; b < kRadixBins; b += blockDim.x) sh[b] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kTile;
#pragma unroll
  for (int r = 0; r < kKeysPerThread; r++) {
    const uint32_t i = base + r * kLbvhBlock + threadIdx.x;
    if (i < n) atomicAdd(&sh[static_cast<uint32_t>(keys[i] >> shift) & (kRadixBins - 1)], 1u);
  }
  __syncthreads();
  for (int b = threadIdx.x; b < kRadixBins; b += blockDim.x) hist[static_cast<size_t>(b) * n_tiles + blockIdx.x] = sh[b];
}

// Exclusive scan of `count` uint32 in place, single block (count = 256 * n_tiles: a few million at most).
__global__ void __launch_bounds__(1024) k_radix_scan(uint32_t* __restrict__ data, uint32_t count) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < count; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < count ? data[i] : 0u;
    uint32_t x = v;
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, off);
      if ((threadIdx.x & 31) >= static_cast<unsigned>(off)) x += y;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t w = warp_sums[threadIdx.x];
      for (int off = 1; off < 32; off <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, w, off);
        if (threadIdx.x >= static_cast<unsigned>(off)) w += y;
      }
      warp_sums[threadIdx.x] = w;
    }
    __syncthreads();
    const uint32_t warp_prefix = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0u;
    const uint32_t incl = x + warp_prefix + carry;
    if (i < count) data[i] = incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
}

// Stable scatter: the rank of a key among equal digits of its tile follows the (round, warp, lane) = global index order.
__global__ void __launch_bounds__(kLbvhBlock) k_radix_scatter(const unsigned long long* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                              uint32_t n, int shift, uint32_t n_tiles, const uint32_t* __restrict__ offsets,
                                                              unsigned long long* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
  constexpr int kWarps = kLbvhBlock / 32;
  __shared__ uint32_t running[kRadixBins];          // keys of this digit already placed by earlier rounds
  __shared__ uint32_t warp_cnt[kWarps][kRadixBins];  // per-round, per-warp digit counts -> exclusive prefix over warps
  __shared__ uint32_t glob[kRadixBins];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (int b = threadIdx.x; b < kRadixBins; b += blockDim.x) {
    running[b] = 0;
    glob[b] = offsets[static_cast<size_t>(b) * n_tiles + blockIdx.x];
  }
  const uint32_t base = blockIdx.x * kTile;
  for (int r = 0; r < kKeysPerThread; r++) {
    for (int b = threadIdx.x; b < kWarps * kRadixBins; b += blockDim.x) (&warp_cnt[0][0])[b] = 0;
    __syncthreads();
    const uint32_t i = base + r * kLbvhBlock + threadIdx.x;
    const bool valid = i < n;
    unsigned long long key = 0;
    uint32_t val = 0, digit = kRadixBins;  // invalid lanes get a digit no valid lane has
    if (valid) {
      key = keys_in[i];
      val = vals_in[i];
      digit = static_cast<uint32_t>(key >> shift) & (kRadixBins - 1);
    }
    const unsigned peers = __match_any_sync(0xFFFFFFFFu, digit);
    const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank_in_warp == 0) warp_cnt[warp][digit] = __popc(peers);
    __syncthreads();
    // exclusive prefix over warps for every digit; thread b owns digit b (blockDim == kRadixBins)
    {
      const int b = threadIdx.x;
      uint32_t acc = 0;
#pragma unroll
      for (int w = 0; w < kWarps; w++) {
        const uint32_t c = warp_cnt[w][b];
        warp_cnt[w][b] = acc;
        acc += c;
      }
      // total of this round is folded into `running` after the keys of the round are placed
      __syncthreads();
      if (valid) {
        const uint32_t pos = glob[digit] + running[digit] + warp_cnt[warp][digit] + rank_in_warp;
        keys_out[pos] = key;
        vals_out[pos] = val;
      }
      __syncthreads();
      running[b] += acc;
    }
    __syncthreads();
  }
}

// ---- hierarchy ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kLbvhBlock) k_lbvh_gather(const rt2::BuildPrim* __restrict__ prims, const uint32_t* __restrict__ sorted_vals, uint32_t n,
                                                            float4* __restrict__ leaf_min, float4* __restrict__ leaf_max,
                                                            uint32_t* __restrict__ prim_refs_out) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const rt2::BuildPrim p = prims[sorted_vals[j]];
  leaf_min[j] = make_float4(p.bmin[0], p.bmin[1], p.bmin[2], 0.0f);
  leaf_max[j] = make_float4(p.bmax[0], p.bmax[1], p.bmax[2], 0.0f);
  prim_refs_out[j] = p.ref;
}

// common-prefix length of keys i and j (index tie-break for duplicate keys), -1 outside [0, n)
__device__ __forceinline__ int lbvh_delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
  if (j < 0 || j >= n) return -1;
  const unsigned long long a = keys[i], b = keys[j];
  if (a == b) return 64 + __clz(static_cast<uint32_t>(i) ^ static_cast<uint32_t>(j));
  return __clzll(static_cast<long long>(a ^ b));
}

constexpr uint32_t kChildLeaf = 0x80000000u;

// Karras 2012, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees", Fig. 4.
__global__ void __launch_bounds__(kLbvhBlock) k_lbvh_hierarchy(const unsigned long long* __restrict__ keys, int n, uint2* __restrict__ children,
                                                               uint32_t* __restrict__ parent_internal, uint32_t* __restrict__ parent_leaf) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
  const int dmin = lbvh_delta(keys, n, i, i - d);
  int lmax = 2;
  while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
  int l = 0;
  for (int t = lmax >> 1; t >= 1; t >>= 1) {
    if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
  }
  const int j = i + l * d;
  const int dnode = lbvh_delta(keys, n, i, j);
  int s = 0;
  for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
    if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    if (t == 1) break;
  }
  const int gamma = i + s * d + min(d, 0);
  const int lo = min(i, j), hi = max(i, j);
  uint32_t left, right;
  if (lo == gamma) {
    left = kChildLeaf | static_cast<uint32_t>(gamma);
    parent_leaf[gamma] = static_cast<uint32_t>(i);
  } else {
    left = static_cast<uint32_t>(gamma);
    parent_internal[gamma] = static_cast<uint32_t>(i);
  }
  if (hi == gamma + 1) {
    right = kChildLeaf | static_cast<uint32_t>(gamma + 1);
    parent_leaf[gamma + 1] = static_cast<uint32_t>(i);
  } else {
    right = static_cast<uint32_t>(gamma + 1);
    parent_internal[gamma + 1] = static_cast<uint32_t>(i);
  }
  children[i] = make_uint2(left, right);
  if (i == 0) parent_internal[0] = 0xFFFFFFFFu;
}

// Bottom-up boxes: every leaf walks towards the root; the first thread to reach a node stops, the second (which sees
// both children complete) forms the union and continues.
__global__ void __launch_bounds__(kLbvhBlock) k_lbvh_refit(int n, const uint2* __restrict__ children, const uint32_t* __restrict__ parent_internal,
                                                           const uint32_t* __restrict__ parent_leaf, const float4* __restrict__ leaf_min,
                                                           const float4* __restrict__ leaf_max, float4* __restrict__ node_min,
                                                           float4* __restrict__ node_max, uint32_t* __restrict__ visit) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  uint32_t node = parent_leaf[j];
  while (node != 0xFFFFFFFFu) {
    __threadfence();
    if (atomicAdd(&visit[node], 1u) == 0u) return;
    const uint2 c = children[node];
    const uint32_t li = c.x & ~kChildLeaf, ri = c.y & ~kChildLeaf;
    const volatile float4* lmin = (c.x & kChildLeaf) ? leaf_min + li : node_min + li;
    const volatile float4* lmax = (c.x & kChildLeaf) ? leaf_max + li : node_max + li;
    const volatile float4* rmin = (c.y & kChildLeaf) ? leaf_min + ri : node_min + ri;
    const volatile float4* rmax = (c.y & kChildLeaf) ? leaf_max + ri : node_max + ri;
    const float4 mn = make_float4(fminf(lmin->x, rmin->x), fminf(lmin->y, rmin->y), fminf(lmin->z, rmin->z), 0.0f);
    const float4 mx = make_float4(fmaxf(lmax->x, rmax->x), fmaxf(lmax->y, rmax->y), fmaxf(lmax->z, rmax->z), 0.0f);
    node_min[node] = mn;
    node_max[node] = mx;
    node = parent_internal[node];
  }
}

// Internal node i -> node pair (pair_base + i): each 32-byte node carries ONE child's box and its link.
__global__ void __launch_bounds__(kLbvhBlock) k_lbvh_emit(int n, uint32_t pair_base, uint32_t ref_base, const uint2* __restrict__ children,
                                                          const float4* __restrict__ leaf_min, const float4* __restrict__ leaf_max,
                                                          const float4* __restrict__ node_min, const float4* __restrict__ node_max,
                                                          const uint32_t* __restrict__ prim_refs, float4* __restrict__ nodes_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n - 1) return;
  const uint2 c = children[i];
  float4* out = nodes_out + (static_cast<size_t>(pair_base) + i) * 4;
#pragma unroll
  for (int side = 0; side < 2; side++) {
    const uint32_t link = side ? c.y : c.x;
    const uint32_t idx = link & ~kChildLeaf;
    float4 mn, mx;
    uint32_t left_first, count;
    if (link & kChildLeaf) {
      mn = leaf_min[idx];
      mx = leaf_max[idx];
      // device node format (rt_trace.cuh make_leaf_entry): a one-primitive leaf carries the primitive reference itself;
      // .w of the max corner keeps the ABI values as count << 27 | left_first
      left_first = kChildLeaf | 0x40000000u | prim_refs[ref_base + idx];
      count = (1u << 27) | (ref_base + idx);
    } else {
      mn = node_min[idx];
      mx = node_max[idx];
      left_first = pair_base + idx;
      count = 0;
    }
    out[2 * side + 0] = make_float4(mn.x, mn.y, mn.z, __uint_as_float(left_first));
    out[2 * side + 1] = make_float4(mx.x, mx.y, mx.z, __uint_as_float(count));
  }
}

// ---- PLOC: parallel locally-ordered clustering (Meister & Bittner 2018) ----------------------------------------------------
// The Karras hierarchy above cuts at Morton-cell boundaries wherever they fall; PLOC builds the tree bottom-up instead: the
// clusters (at first the leaves, in Morton order) each look for the neighbour within +-kPlocRadius positions whose union with
// them has the smallest surface area, mutual nearest neighbours merge into a new node, the survivors are compacted, and the
// loop repeats until one cluster is left — an agglomerative build whose quality is close to a full SAH sweep (measured on the
// stress scene: node pairs per ray, profiles/r02_notes.md).  Node ids are handed out from n - 2 downwards, so the last merge —
// the root — is node 0, as k_lbvh_emit and the traversal expect.
constexpr int kPlocRadius = 8;

__device__ __forceinline__ float box_area(float4 mn, float4 mx) {
  const float dx = mx.x - mn.x, dy = mx.y - mn.y, dz = mx.z - mn.z;
  return dx * dy + dy * dz + dz * dx;
}

__global__ void __launch_bounds__(kLbvhBlock) k_ploc_init(uint32_t n, const float4* __restrict__ leaf_min, const float4* __restrict__ leaf_max,
                                                          float4* __restrict__ cmin, float4* __restrict__ cmax, uint32_t* __restrict__ cnode) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  cmin[i] = leaf_min[i];
  cmax[i] = leaf_max[i];
  cnode[i] = kChildLeaf | i;
}

// nearest neighbour of every cluster inside the window (smallest union area; ties go to the lower index, so that the pair with
// the globally smallest union is always mutual and every round merges at least one pair)
__global__ void __launch_bounds__(kLbvhBlock) k_ploc_nn(uint32_t m, const float4* __restrict__ cmin, const float4* __restrict__ cmax,
                                                        uint32_t* __restrict__ nn) {
  __shared__ float4 smin[kLbvhBlock + 2 * kPlocRadius], smax[kLbvhBlock + 2 * kPlocRadius];
  const int base = static_cast<int>(blockIdx.x * blockDim.x) - kPlocRadius;
  for (int t = threadIdx.x; t < kLbvhBlock + 2 * kPlocRadius; t += blockDim.x) {
    const int g = base + t;
    if (g >= 0 && g < static_cast<int>(m)) {
      smin[t] = cmin[g];
      smax[t] = cmax[g];
    }
  }
  __syncthreads();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const int li = threadIdx.x + kPlocRadius;
  const float4 amin = smin[li], amax = smax[li];
  float best = 3.4e38f;
  uint32_t best_j = i;
  for (int dlt = -kPlocRadius; dlt <= kPlocRadius; dlt++) {
    const int g = static_cast<int>(i) + dlt;
    if (dlt == 0 || g < 0 || g >= static_cast<int>(m)) continue;
    const float4 bmin = smin[li + dlt], bmax = smax[li + dlt];
    const float4 mn = make_float4(fminf(amin.x, bmin.x), fminf(amin.y, bmin.y), fminf(amin.z, bmin.z), 0.0f);
    const float4 mx = make_float4(fmaxf(amax.x, bmax.x), fmaxf(amax.y, bmax.y), fmaxf(amax.z, bmax.z), 0.0f);
    const float ar = box_area(mn, mx);
    if (ar < best) {  // ascending g: the first of equal areas (the lower index) wins
      best = ar;
      best_j = static_cast<uint32_t>(g);
    }
  }
  nn[i] = best_j;
}

// mutual nearest neighbours merge (the lower slot keeps the merged cluster, the upper one is dropped); flag = survives
__global__ void __launch_bounds__(kLbvhBlock) k_ploc_merge(uint32_t m, uint32_t n, float4* __restrict__ cmin, float4* __restrict__ cmax,
                                                           uint32_t* __restrict__ cnode, const uint32_t* __restrict__ nn,
                                                           uint32_t* __restrict__ merged_count, uint2* __restrict__ children,
                                                           uint32_t* __restrict__ parent_internal, uint32_t* __restrict__ parent_leaf,
                                                           float4* __restrict__ node_min, float4* __restrict__ node_max,
                                                           uint32_t* __restrict__ flag) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint32_t j = nn[i];
  const bool mutual = (j != i) && (nn[j] == i);
  if (!mutual) {
    flag[i] = 1u;
    return;
  }
  if (i > j) {
    flag[i] = 0u;
    return;
  }
  const uint32_t id = (n - 2u) - atomicAdd(merged_count, 1u);
  const uint32_t a = cnode[i], b = cnode[j];
  children[id] = make_uint2(a, b);
  if (a & kChildLeaf) parent_leaf[a & ~kChildLeaf] = id;
  else parent_internal[a] = id;
  if (b & kChildLeaf) parent_leaf[b & ~kChildLeaf] = id;
  else parent_internal[b] = id;
  const float4 amin = cmin[i], amax = cmax[i], bmin = cmin[j], bmax = cmax[j];
  const float4 mn = make_float4(fminf(amin.x, bmin.x), fminf(amin.y, bmin.y), fminf(amin.z, bmin.z), 0.0f);
  const float4 mx = make_float4(fmaxf(amax.x, bmax.x), fmaxf(amax.y, bmax.y), fmaxf(amax.z, bmax.z), 0.0f);
  node_min[id] = mn;
  node_max[id] = mx;
  cmin[i] = mn;
  cmax[i] = mx;
  cnode[i] = id;
  flag[i] = 1u;
  if (id == 0u) parent_internal[0] = 0xFFFFFFFFu;  // the last merge is the root
}

// survivors move to their rank (pos = exclusive prefix sum of the flags)
__global__ void __launch_bounds__(kLbvhBlock) k_ploc_compact(uint32_t m, const uint32_t* __restrict__ pos, const uint32_t* __restrict__ nn_mutual,
                                                             const float4* __restrict__ cmin, const float4* __restrict__ cmax,
                                                             const uint32_t* __restrict__ cnode, float4* __restrict__ omin,
                                                             float4* __restrict__ omax, uint32_t* __restrict__ onode, uint32_t* __restrict__ new_count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  const uint32_t p = pos[i];
  const bool alive = (i + 1 < m) ? (pos[i + 1] != p) : (nn_mutual[0] != 0u);  // nn_mutual[0] = flag of the last element (saved before the scan)
  if (alive) {
    omin[p] = cmin[i];
    omax[p] = cmax[i];
    onode[p] = cnode[i];
  }
  if (i + 1 == m) *new_count = p + (alive ? 1u : 0u);
}

// Depth of the tree in node pairs (= the number of stack entries a traversal may need): every leaf walks to the root.
__global__ void __launch_bounds__(kLbvhBlock) k_lbvh_depth(int n, const uint32_t* __restrict__ parent_internal,
                                                           const uint32_t* __restrict__ parent_leaf, uint32_t* __restrict__ depth_out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t depth = 0;
  if (j < n) {
    uint32_t node = parent_leaf[j];
    depth = 1;
    while ((node = parent_internal[node]) != 0xFFFFFFFFu) depth++;
  }
  for (int off = 16; off > 0; off >>= 1) depth = max(depth, __shfl_down_sync(0xFFFFFFFFu, depth, off));
  if ((threadIdx.x & 31) == 0 && depth) atomicMax(depth_out, depth);
}

// Trees with 0 or 1 primitive: {leaf | empty, empty}
__global__ void k_lbvh_tiny(int n, uint32_t pair_base, uint32_t ref_base, const rt2::BuildPrim* __restrict__ prims, float4* __restrict__ nodes_out,
                            uint32_t* __restrict__ prim_refs_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float nanv = __int_as_float(0x7FC00000);
  float4* out = nodes_out + static_cast<size_t>(pair_base) * 4;
  for (int k = 0; k < 4; k++) out[k] = make_float4(nanv, nanv, nanv, __uint_as_float(0u));
  if (n == 1) {
    const rt2::BuildPrim p = prims[0];
    out[0] = make_float4(p.bmin[0], p.bmin[1], p.bmin[2], __uint_as_float(kChildLeaf | 0x40000000u | p.ref));
    out[1] = make_float4(p.bmax[0], p.bmax[1], p.bmax[2], __uint_as_float((1u << 27) | ref_base));
    prim_refs_out[0] = p.ref;
  }
}

}  // namespace rt2dev

namespace rt2 {

using namespace rt2dev;

#define LBVH_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      *err = std::string(#call) + " failed: " + cudaGetErrorString(e__);            \
      return RT2_ERR_CUDA;                                                           \
    }                                                                                \
  } while (0)

int BuildLbvhOnDevice(const BuildPrim* d_prims, uint32_t n, uint32_t pair_base, uint32_t ref_base, void* d_nodes, uint32_t* d_prim_refs,
                      LbvhScratch* scratch, void* stream_v, uint64_t* launches, uint32_t* d_depth, bool use_ploc, std::string* err) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  float4* nodes = static_cast<float4*>(d_nodes);
  if (n <= 1) {
    k_lbvh_tiny<<<1, 32, 0, stream>>>(static_cast<int>(n), pair_base, ref_base, d_prims, nodes, d_prim_refs + ref_base);
    (*launches)++;
    const uint32_t one = 1;
    LBVH_CUDA(cudaMemcpyAsync(d_depth, &one, sizeof(one), cudaMemcpyHostToDevice, stream));
    LBVH_CUDA(cudaStreamSynchronize(stream));  // `one` is a stack variable
    return RT2_OK;
  }
  // scratch (grown on demand, reused between trees); every sub-buffer is 256-byte aligned
  const uint32_t n_tiles = (n + kTile - 1) / kTile;
  auto align = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
  const size_t sz_u32 = align(n * 4ull), sz_f4 = align(n * 16ull), sz_u2 = align(n * 8ull);
  const size_t sz_hist = align(static_cast<size_t>(kRadixBins) * n_tiles * 4);
  const size_t need = 256 + 2 * sz_u2 + 2 * sz_u32 + sz_hist + 4 * sz_f4 + sz_u2 + 3 * sz_u32 + (use_ploc ? 4 * sz_f4 : 0);
  if (need > scratch->bytes) {
    if (scratch->ptr) cudaFree(scratch->ptr);
    scratch->ptr = nullptr;
    scratch->bytes = 0;
    LBVH_CUDA(cudaMalloc(&scratch->ptr, need));
    scratch->bytes = need;
  }
  char* p = static_cast<char*>(scratch->ptr);
  auto take = [&](size_t bytes) {
    char* r = p;
    p += bytes;
    return r;
  };
  char* head = take(256);
  int* bounds = reinterpret_cast<int*>(head);                // 6 ints
  double* moments = reinterpret_cast<double*>(head + 64);   // 6 doubles
  float* grid = reinterpret_cast<float*>(head + 128);        // 6 floats
  unsigned long long* key_or = reinterpret_cast<unsigned long long*>(head + 192);
  unsigned long long* keys_a = reinterpret_cast<unsigned long long*>(take(sz_u2));
  unsigned long long* keys_b = reinterpret_cast<unsigned long long*>(take(sz_u2));
  uint32_t* vals_a = reinterpret_cast<uint32_t*>(take(sz_u32));
  uint32_t* vals_b = reinterpret_cast<uint32_t*>(take(sz_u32));
  uint32_t* hist = reinterpret_cast<uint32_t*>(take(sz_hist));
  float4* leaf_min = reinterpret_cast<float4*>(take(sz_f4));
  float4* leaf_max = reinterpret_cast<float4*>(take(sz_f4));
  float4* node_min = reinterpret_cast<float4*>(take(sz_f4));
  float4* node_max = reinterpret_cast<float4*>(take(sz_f4));
  uint2* children = reinterpret_cast<uint2*>(take(sz_u2));
  uint32_t* parent_internal = reinterpret_cast<uint32_t*>(take(sz_u32));
  uint32_t* parent_leaf = reinterpret_cast<uint32_t*>(take(sz_u32));
  uint32_t* visit = reinterpret_cast<uint32_t*>(take(sz_u32));

  auto ordered = [](float f) {
    int i;
    std::memcpy(&i, &f, sizeof(i));
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
  };
  const int init_bounds[6] = {ordered(3.4e38f), ordered(3.4e38f), ordered(3.4e38f), ordered(-3.4e38f), ordered(-3.4e38f), ordered(-3.4e38f)};
  LBVH_CUDA(cudaMemsetAsync(head, 0, 256, stream));
  LBVH_CUDA(cudaMemcpyAsync(bounds, init_bounds, sizeof(init_bounds), cudaMemcpyHostToDevice, stream));
  const uint32_t grid_n = (n + kLbvhBlock - 1) / kLbvhBlock;
  const uint32_t grid_red = grid_n < 1184u ? grid_n : 1184u;
  k_lbvh_bounds<<<grid_red, kLbvhBlock, 0, stream>>>(d_prims, n, bounds, moments);
  k_lbvh_grid<<<1, 32, 0, stream>>>(n, bounds, moments, grid);
  k_lbvh_morton<<<grid_n, kLbvhBlock, 0, stream>>>(d_prims, n, grid, keys_a, vals_a, key_or);
  *launches += 3;
  // which 8-bit digits are non-zero in at least one key?  (a flat or small scene uses far fewer than 63 bits)
  unsigned long long used = 0;
  LBVH_CUDA(cudaMemcpyAsync(&used, key_or, sizeof(used), cudaMemcpyDeviceToHost, stream));
  LBVH_CUDA(cudaStreamSynchronize(stream));
  unsigned long long *kin = keys_a, *kout = keys_b;
  uint32_t *vin = vals_a, *vout = vals_b;
  for (int pass = 0; pass < 8; pass++) {
    const int shift = pass * kRadixBits;
    if (((used >> shift) & 0xFFull) == 0ull) continue;  // every key has a zero digit here: the pass would be the identity
    k_radix_hist<<<n_tiles, kLbvhBlock, 0, stream>>>(kin, n, shift, n_tiles, hist);
    k_radix_scan<<<1, 1024, 0, stream>>>(hist, static_cast<uint32_t>(kRadixBins) * n_tiles);
    k_radix_scatter<<<n_tiles, kLbvhBlock, 0, stream>>>(kin, vin, n, shift, n_tiles, hist, kout, vout);
    *launches += 3;
    unsigned long long* tk = kin;
    kin = kout;
    kout = tk;
    uint32_t* tv = vin;
    vin = vout;
    vout = tv;
  }
  // the sorted data is in (kin, vin)
  k_lbvh_gather<<<grid_n, kLbvhBlock, 0, stream>>>(d_prims, vin, n, leaf_min, leaf_max, d_prim_refs + ref_base);
  if (use_ploc) {
    // PLOC: the radix-sort buffers are free again and hold the cluster arrays (boxes double-buffered for the compaction)
    float4* cmin[2] = {reinterpret_cast<float4*>(take(sz_f4)), reinterpret_cast<float4*>(take(sz_f4))};
    float4* cmax[2] = {reinterpret_cast<float4*>(take(sz_f4)), reinterpret_cast<float4*>(take(sz_f4))};
    uint32_t* cnode[2] = {vals_a, vals_b};
    uint32_t* nn = reinterpret_cast<uint32_t*>(keys_a);
    uint32_t* flag = reinterpret_cast<uint32_t*>(keys_b);
    uint32_t* counters = reinterpret_cast<uint32_t*>(head + 200);  // [0] merges so far, [1] clusters after compaction, [2] last flag
    k_ploc_init<<<grid_n, kLbvhBlock, 0, stream>>>(n, leaf_min, leaf_max, cmin[0], cmax[0], cnode[0]);
    (*launches)++;
    uint32_t m = n;
    int cur = 0;
    while (m > 1) {
      const uint32_t gm = (m + kLbvhBlock - 1) / kLbvhBlock;
      k_ploc_nn<<<gm, kLbvhBlock, 0, stream>>>(m, cmin[cur], cmax[cur], nn);
      k_ploc_merge<<<gm, kLbvhBlock, 0, stream>>>(m, n, cmin[cur], cmax[cur], cnode[cur], nn, counters, children, parent_internal, parent_leaf,
                                                  node_min, node_max, flag);
      LBVH_CUDA(cudaMemcpyAsync(counters + 2, flag + (m - 1), sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
      k_radix_scan<<<1, 1024, 0, stream>>>(flag, m);
      k_ploc_compact<<<gm, kLbvhBlock, 0, stream>>>(m, flag, counters + 2, cmin[cur], cmax[cur], cnode[cur], cmin[cur ^ 1], cmax[cur ^ 1],
                                                    cnode[cur ^ 1], counters + 1);
      *launches += 4;
      uint32_t m_new = 0;
      LBVH_CUDA(cudaMemcpyAsync(&m_new, counters + 1, sizeof(m_new), cudaMemcpyDeviceToHost, stream));
      LBVH_CUDA(cudaStreamSynchronize(stream));
      if (m_new == 0 || m_new >= m) {
        *err = "PLOC made no progress (" + std::to_string(m) + " -> " + std::to_string(m_new) + " clusters)";
        return RT2_ERR_STATE;
      }
      m = m_new;
      cur ^= 1;
    }
  } else {
    LBVH_CUDA(cudaMemsetAsync(visit, 0, n * 4ull, stream));
    k_lbvh_hierarchy<<<grid_n, kLbvhBlock, 0, stream>>>(kin, static_cast<int>(n), children, parent_internal, parent_leaf);
    k_lbvh_refit<<<grid_n, kLbvhBlock, 0, stream>>>(static_cast<int>(n), children, parent_internal, parent_leaf, leaf_min, leaf_max, node_min,
                                                    node_max, visit);
    *launches += 2;
  }
  k_lbvh_emit<<<grid_n, kLbvhBlock, 0, stream>>>(static_cast<int>(n), pair_base, ref_base, children, leaf_min, leaf_max, node_min, node_max, d_prim_refs,
                                                 nodes);
  k_lbvh_depth<<<grid_n, kLbvhBlock, 0, stream>>>(static_cast<int>(n), parent_internal, parent_leaf, d_depth);
  *launches += 3;
  LBVH_CUDA(cudaGetLastError());
  return RT2_OK;
}

void FreeLbvhScratch(LbvhScratch* s) {
  if (s->ptr) cudaFree(s->ptr);
  s->ptr = nullptr;
  s->bytes = 0;
}

}  // namespace rt2
