// Ray reordering for the extend stage: a stable LSD radix sort of (coherence key -> queue index) whose element count
// lives in device memory, so a whole batch is enqueued without host synchronisation.
//
// Why: after the first bounce the queue order (material bin, then push order) says little about where a ray starts or
// where it is heading; 32 unrelated rays per warp make k_traverse wait for the longest descent in every phase
// (profiles/r01_notes.md: 9 of 32 lanes active per instruction).  Rays that share an origin cell and a direction octant
// walk the same sub-trees in the same front-to-back order, so their descent lengths — and node fetches — agree.
// The reference has no counterpart (it traces one recursive path per thread, RayTracer.cpp:55-70); the reordering does
// not change any result: k_traverse writes trav[ray index], every later stage reads by ray index.
//
// Key (kSortKeyBits = 18): [17:15] direction octant, [14:0] 15-bit Morton code of the origin in a 32^3 grid over the
// robust scene bounds (DeviceScene::sort_lo / sort_scale).  Two passes of 9 bits:
//   k_sort_hist     per-block digit histogram of the block's contiguous chunk        R 4 B / ray
//   k_sort_scan     exclusive offsets: over blocks within a digit, then over digits
//   k_sort_scatter  stable scatter, ranks by __match_any_sync + per-warp counters     R 8 + W 8 B / ray
// Queues below kSortMinRays skip every kernel (the count is checked on the device) and are traversed in queue order.
#pragma once
#include "rt_trace.cuh"

namespace rt2dev {

constexpr int kSortBlock = 256;
constexpr int kSortWarps = kSortBlock / 32;
constexpr int kSortDigitBits = 9;
constexpr int kSortBins = 1 << kSortDigitBits;
constexpr int kSortKeysPerThread = 8;
constexpr int kSortTile = kSortBlock * kSortKeysPerThread;  // 2048 keys per tile
constexpr int kSortKeyBits = 18;

__device__ __forceinline__ uint32_t spread_bits5(uint32_t v) {
  // abcde -> a00b00c00d00e
  v = (v | (v << 8)) & 0x0000100Fu;
  v = (v | (v << 4)) & 0x000010C3u;
  v = (v | (v << 2)) & 0x00001249u;
  return v;
}

// Coherence key of a ray (see the header comment).  Not part of any result: only the traversal ORDER depends on it.
__device__ __forceinline__ uint32_t ray_sort_key(const DeviceScene& S, F3 o, F3 d) {
  const float fx = fminf(fmaxf((o.x - S.sort_lo[0]) * S.sort_scale[0], 0.0f), 31.0f);
  const float fy = fminf(fmaxf((o.y - S.sort_lo[1]) * S.sort_scale[1], 0.0f), 31.0f);
  const float fz = fminf(fmaxf((o.z - S.sort_lo[2]) * S.sort_scale[2], 0.0f), 31.0f);
  const uint32_t m = (spread_bits5(static_cast<uint32_t>(fx)) << 2) | (spread_bits5(static_cast<uint32_t>(fy)) << 1) |
                     spread_bits5(static_cast<uint32_t>(fz));
  const uint32_t oct = (d.x < 0.0f ? 4u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 1u : 0u);
  return (oct << 15) | m;
}

// Contiguous chunk of the queue owned by a block: whole tiles, so that chunk borders never split a warp's 256-key run.
__device__ __forceinline__ void sort_chunk(uint32_t n, uint32_t& begin, uint32_t& end) {
  const uint32_t tiles = (n + kSortTile - 1) / kSortTile;
  const uint32_t per_block = (tiles + gridDim.x - 1) / gridDim.x;
  const uint64_t b = static_cast<uint64_t>(blockIdx.x) * per_block * kSortTile;
  const uint64_t e = b + static_cast<uint64_t>(per_block) * kSortTile;
  begin = b < n ? static_cast<uint32_t>(b) : n;
  end = e < n ? static_cast<uint32_t>(e) : n;
}

// hist[block * kSortBins + digit]
__global__ void __launch_bounds__(kSortBlock) k_sort_hist(const uint32_t* __restrict__ counters, uint32_t min_rays,
                                                          const uint32_t* __restrict__ keys, int shift, uint32_t* __restrict__ hist) {
  const uint32_t n = counters[0];
  if (n < min_rays) return;
  __shared__ uint32_t sh[kSortBins];
  for (int b = threadIdx.x; b < kSortBins; b += kSortBlock) sh[b] = 0;
  __syncthreads();
  uint32_t begin, end;
  sort_chunk(n, begin, end);
  for (uint32_t i = begin + threadIdx.x; i < end; i += kSortBlock) atomicAdd(&sh[(keys[i] >> shift) & (kSortBins - 1)], 1u);
  __syncthreads();
  for (int b = threadIdx.x; b < kSortBins; b += kSortBlock) hist[static_cast<size_t>(blockIdx.x) * kSortBins + b] = sh[b];
}

// One block, one thread per digit: running sum over the blocks (in place -> exclusive offset of the block within the
// digit), then an exclusive scan of the digit totals into bin_base.
__global__ void __launch_bounds__(kSortBins) k_sort_scan(const uint32_t* __restrict__ counters, uint32_t min_rays, uint32_t n_blocks,
                                                         uint32_t* __restrict__ hist, uint32_t* __restrict__ bin_base) {
  if (counters[0] < min_rays) return;
  const uint32_t bin = threadIdx.x;
  uint32_t run = 0;
  uint32_t b = 0;
  for (; b + 8 <= n_blocks; b += 8) {
    uint32_t v[8];
#pragma unroll
    for (int k = 0; k < 8; k++) v[k] = hist[static_cast<size_t>(b + k) * kSortBins + bin];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      hist[static_cast<size_t>(b + k) * kSortBins + bin] = run;
      run += v[k];
    }
  }
  for (; b < n_blocks; b++) {
    const uint32_t v = hist[static_cast<size_t>(b) * kSortBins + bin];
    hist[static_cast<size_t>(b) * kSortBins + bin] = run;
    run += v;
  }
  // exclusive scan of `run` over the kSortBins threads of the block
  __shared__ uint32_t warp_sums[kSortBins / 32];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t x = run;
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, off);
    if (lane >= static_cast<unsigned>(off)) x += y;
  }
  if (lane == 31u) warp_sums[warp] = x;
  __syncthreads();
  if (warp == 0 && lane < kSortBins / 32) {
    uint32_t w = warp_sums[lane];
    for (int off = 1; off < kSortBins / 32; off <<= 1) {
      const uint32_t y = __shfl_up_sync(0x0000FFFFu, w, off);
      if (lane >= static_cast<unsigned>(off)) w += y;
    }
    warp_sums[lane] = w;
  }
  __syncthreads();
  bin_base[bin] = x - run + (warp ? warp_sums[warp - 1] : 0u);
}

// Stable scatter of the block's chunk.  vals_in == nullptr: the value of key i is i (first pass); keys_out == nullptr:
// keys are not needed any more (last pass).
__global__ void __launch_bounds__(kSortBlock) k_sort_scatter(const uint32_t* __restrict__ counters, uint32_t min_rays,
                                                             const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                             int shift, const uint32_t* __restrict__ hist,
                                                             const uint32_t* __restrict__ bin_base, uint32_t* __restrict__ keys_out,
                                                             uint32_t* __restrict__ vals_out) {
  const uint32_t n = counters[0];
  if (n < min_rays) return;
  __shared__ uint32_t running[kSortBins];             // next free position of each digit for this block
  __shared__ uint32_t warp_cnt[kSortWarps][kSortBins];  // per-tile digit counts per warp -> start position per warp
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  for (int b = threadIdx.x; b < kSortBins; b += kSortBlock) running[b] = bin_base[b] + hist[static_cast<size_t>(blockIdx.x) * kSortBins + b];
  uint32_t begin, end;
  sort_chunk(n, begin, end);
  for (uint32_t tile = begin; tile < end; tile += kSortTile) {
    for (int b = threadIdx.x; b < kSortWarps * kSortBins; b += kSortBlock) (&warp_cnt[0][0])[b] = 0;
    __syncthreads();
    uint32_t key[kSortKeysPerThread], val[kSortKeysPerThread], rank[kSortKeysPerThread];
    // warp w owns keys [tile + 256 w, tile + 256 (w+1)) in 8 rounds of 32: rank order = index order (stability)
#pragma unroll
    for (int r = 0; r < kSortKeysPerThread; r++) {
      const uint32_t i = tile + warp * (32 * kSortKeysPerThread) + r * 32 + lane;
      const bool valid = i < end;
      key[r] = valid ? keys_in[i] : 0u;
      val[r] = valid ? (vals_in ? vals_in[i] : i) : 0xFFFFFFFFu;
    }
#pragma unroll
    for (int r = 0; r < kSortKeysPerThread; r++) {
      const bool valid = val[r] != 0xFFFFFFFFu;
      const uint32_t digit = valid ? ((key[r] >> shift) & (kSortBins - 1)) : kSortBins;  // invalid lanes match only each other
      const unsigned peers = __match_any_sync(0xFFFFFFFFu, digit);
      const uint32_t before = __popc(peers & lt);
      uint32_t start = 0;
      if (valid) start = warp_cnt[warp][digit];
      __syncwarp();
      if (valid && before == 0) warp_cnt[warp][digit] = start + __popc(peers);
      __syncwarp();
      rank[r] = start + before;
    }
    __syncthreads();
    // exclusive prefix over the warps for every digit, offset by the block's running position
    for (int b = threadIdx.x; b < kSortBins; b += kSortBlock) {
      uint32_t acc = running[b];
#pragma unroll
      for (int w = 0; w < kSortWarps; w++) {
        const uint32_t c = warp_cnt[w][b];
        warp_cnt[w][b] = acc;
        acc += c;
      }
      running[b] = acc;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kSortKeysPerThread; r++) {
      if (val[r] != 0xFFFFFFFFu) {
        const uint32_t digit = (key[r] >> shift) & (kSortBins - 1);
        const uint32_t pos = warp_cnt[warp][digit] + rank[r];
        if (keys_out) keys_out[pos] = key[r];
        vals_out[pos] = val[r];
      }
    }
    __syncthreads();
  }
}

}  // namespace rt2dev
