// 32-byte node pairs: the traversal kernel is bound by the L1 data path (every lane of a warp reads its own node pair, one
// wavefront per 32-byte sector and lane; profiles/r02_notes.md), so the pair is halved from 64 to 32 bytes.  Both child boxes
// are stored as 15-bit coordinates on ONE grid per scene, each in a 16-bit field whose top bit is set: a single PRMT turns a
// field into the float 1 + q / 32768 (bytes {00, lo, hi, 3F}), and the grid's origin and cell size are folded into the ray's
// slab coefficients once per ray — the node step costs one PRMT per coordinate and no conversion instruction.
//   pair = { uint4 child0, uint4 child1 },  child = { minx | maxx << 16, miny | maxy << 16, minz | maxz << 16, entry }
// The min and max plane of an axis share a word, so a per-ray PRMT selector (low or high half, from the sign of the direction)
// yields the plane the ray enters through or the one it leaves through directly: no min / max per axis in the slab test.
// Boxes are rounded outwards and padded by one cell, so the quantised box contains the float box with a margin far above the
// rounding error of the folded slab test: closest hits cannot change (box tests only cull).  Whether a scene
// uses these nodes is decided from the measured surface-area inflation (Renderer::UploadScene).
// The pairs keep the builder's (depth-first) order: laying the top 6 / 10 / 14 / all levels out breadth first, so that the nodes
// every ray visits share cache lines, measured 6 557 / 6 560 / 6 550 / 6 549 against 6 545 Mrays/s — L1 capacity is not the limit.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "../../../include/rt2.h"
#include "rt_qnodes.hpp"

namespace rt2dev {

constexpr int kQBlock = 256;

struct GridParams {
  float lo[3];        // grid origin
  float inv_cell[3];  // 32768 / extent
};

__device__ __forceinline__ bool box_valid(const float4 mn, const float4 mx) {
  return mn.x <= mx.x && mn.y <= mx.y && mn.z <= mx.z;  // false for the NaN bounds of an empty slot
}

__device__ __forceinline__ float half_area(const float ex, const float ey, const float ez) { return ex * ey + ey * ez + ez * ex; }

// out[0]: number of pairs with two empty slots; out[1]: pairs with a non-finite bound; sums[0]: number of child boxes inside the
// grid with a non-zero surface area, sums[1]: sum over those boxes of (quantised area / float area).  The UNWEIGHTED mean is the
// measure: weighting by area lets one giant box (a ground sphere of radius 1000) hide that every small leaf doubled.
__global__ void __launch_bounds__(kQBlock) k_quantise_nodes(const float4* __restrict__ nodes, uint32_t n_pairs, const GridParams g,
                                                            uint4* __restrict__ qnodes, uint32_t* __restrict__ out,
                                                            double* __restrict__ sums) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  double n_boxes = 0.0, ratio_sum = 0.0;
  if (p < n_pairs) {
    float4 mn[2] = {nodes[4ull * p + 0], nodes[4ull * p + 2]};
    float4 mx[2] = {nodes[4ull * p + 1], nodes[4ull * p + 3]};
    const bool v0 = box_valid(mn[0], mx[0]), v1 = box_valid(mn[1], mx[1]);
    if (!v0 && !v1) atomicAdd(&out[0], 1u);
    // an empty slot becomes a copy of its sibling: testing a leaf twice (or walking a subtree twice) cannot change the closest hit
    if (!v0) mn[0] = mn[1], mx[0] = mx[1];
    if (!v1) mn[1] = mn[0], mx[1] = mx[0];
    bool finite = true;
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const float lo3[3] = {mn[c].x, mn[c].y, mn[c].z}, hi3[3] = {mx[c].x, mx[c].y, mx[c].z};
      uint32_t qlo[3], qhi[3];
      float eq[3];
      bool clamped = false;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        finite = finite && isfinite(lo3[k]) && isfinite(hi3[k]);
        // outward rounding + one cell of padding; the clamp only matters for trees the grid does not span (never walked)
        const float a = floorf((lo3[k] - g.lo[k]) * g.inv_cell[k]) - 1.0f;
        const float b = ceilf((hi3[k] - g.lo[k]) * g.inv_cell[k]) + 1.0f;
        clamped = clamped || !(a >= 0.0f && b <= 32767.0f);
        qlo[k] = static_cast<uint32_t>(fminf(fmaxf(a, 0.0f), 32767.0f));
        qhi[k] = static_cast<uint32_t>(fminf(fmaxf(b, 0.0f), 32767.0f));
        eq[k] = static_cast<float>(qhi[k] - qlo[k]) / g.inv_cell[k];
      }
      const float area = half_area(hi3[0] - lo3[0], hi3[1] - lo3[1], hi3[2] - lo3[2]);
      if ((v0 || v1) && !clamped && area > 0.0f) {
        n_boxes += 1.0;
        ratio_sum += static_cast<double>(half_area(eq[0], eq[1], eq[2]) / area);
      }
      const uint32_t kTop = 0x8000u;
      qnodes[2ull * p + c] = make_uint4((kTop | qlo[0]) | ((kTop | qhi[0]) << 16), (kTop | qlo[1]) | ((kTop | qhi[1]) << 16),
                                        (kTop | qlo[2]) | ((kTop | qhi[2]) << 16), __float_as_uint(mn[c].w));
    }
    if (!finite && (v0 || v1)) atomicAdd(&out[1], 1u);
  }
  for (int off = 16; off > 0; off >>= 1) {
    n_boxes += __shfl_down_sync(0xFFFFFFFFu, n_boxes, off);
    ratio_sum += __shfl_down_sync(0xFFFFFFFFu, ratio_sum, off);
  }
  if ((threadIdx.x & 31u) == 0u && n_boxes > 0.0) {
    atomicAdd(&sums[0], n_boxes);
    atomicAdd(&sums[1], ratio_sum);
  }
}

}  // namespace rt2dev

namespace rt2 {

#define QN_CUDA(call)                                                               \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess) {                                                        \
      *err = std::string(#call) + " failed: " + cudaGetErrorString(e_);            \
      return RT2_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

int QuantiseNodesOnDevice(const void* d_nodes, uint32_t n_pairs, const uint32_t* roots, uint32_t n_roots, void* d_qnodes, NodeGrid* grid,
                          void* stream_v, uint64_t* launches, std::string* err) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_v);
  grid->usable = false;
  grid->inflation = 0.0;
  if (n_pairs == 0 || n_roots == 0) return RT2_OK;
  // the grid spans the child boxes of the root pairs (they bound their whole trees)
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  QN_CUDA(cudaStreamSynchronize(stream));
  for (uint32_t r = 0; r < n_roots; r++) {
    if (roots[r] >= n_pairs) {
      *err = "tree root outside the node array";
      return RT2_ERR_STATE;
    }
    float pr[16];
    QN_CUDA(cudaMemcpy(pr, static_cast<const char*>(d_nodes) + 64ull * roots[r], sizeof(pr), cudaMemcpyDeviceToHost));
    for (int c = 0; c < 2; c++) {
      const float* mn = pr + 8 * c;
      const float* mx = pr + 8 * c + 4;
      if (!(mn[0] <= mx[0] && mn[1] <= mx[1] && mn[2] <= mx[2])) continue;  // empty slot
      for (int k = 0; k < 3; k++) {
        lo[k] = std::fmin(lo[k], mn[k]);
        hi[k] = std::fmax(hi[k], mx[k]);
      }
    }
  }
  float span_max = 0.0f;
  for (int k = 0; k < 3; k++) {
    if (!std::isfinite(lo[k]) || !std::isfinite(hi[k])) return RT2_OK;  // empty or unbounded scene: float nodes
    span_max = std::fmax(span_max, hi[k] - lo[k]);
  }
  if (!(span_max > 0.0f) || !std::isfinite(span_max)) return RT2_OK;
  rt2dev::GridParams g;
  for (int k = 0; k < 3; k++) {
    // 16 cells of margin below, ~48 above; a flat axis still gets a non-zero cell
    const float span = std::fmax(hi[k] - lo[k], span_max * 1e-6f);
    const float ext = span * (1.0f + 1.0f / 512.0f);
    g.lo[k] = lo[k] - span * (1.0f / 2048.0f);
    g.inv_cell[k] = 32768.0f / ext;
    grid->ext[k] = ext;
    grid->base[k] = g.lo[k] - ext;
    if (!std::isfinite(g.inv_cell[k]) || !std::isfinite(grid->base[k])) return RT2_OK;
  }
  uint32_t* d_out = nullptr;
  QN_CUDA(cudaMalloc(&d_out, 2 * sizeof(uint32_t) + 2 * sizeof(double) + 8));
  double* d_sums = reinterpret_cast<double*>(d_out + 2);
  QN_CUDA(cudaMemsetAsync(d_out, 0, 2 * sizeof(uint32_t) + 2 * sizeof(double), stream));
  rt2dev::k_quantise_nodes<<<(n_pairs + rt2dev::kQBlock - 1) / rt2dev::kQBlock, rt2dev::kQBlock, 0, stream>>>(
      static_cast<const float4*>(d_nodes), n_pairs, g, static_cast<uint4*>(d_qnodes), d_out, d_sums);
  (*launches)++;
  struct {
    uint32_t out[2];
    double sums[2];
  } h;
  cudaError_t e = cudaMemcpyAsync(&h, d_out, sizeof(h), cudaMemcpyDeviceToHost, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  cudaFree(d_out);
  if (e != cudaSuccess) {
    *err = std::string("node quantisation failed: ") + cudaGetErrorString(e);
    return RT2_ERR_CUDA;
  }
  grid->inflation = h.sums[0] > 0.0 ? h.sums[1] / h.sums[0] - 1.0 : 0.0;
  grid->usable = h.out[0] == 0u && h.out[1] == 0u;
  return RT2_OK;
}

}  // namespace rt2
