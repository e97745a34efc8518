// GPU LBVH builder interface (implementation: rt_lbvh.cu).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

#include "../host/scene_host.hpp"

namespace rt2 {

struct LbvhScratch {
  void* ptr{nullptr};
  size_t bytes{0};
};

// Builds one tree over the n device-resident (box, ref) records `d_prims` into node pairs [pair_base, pair_base + max(1, n-1))
// of `d_nodes` and primitive references [ref_base, ref_base + n) of `d_prim_refs`.  Asynchronous on `stream` (apart from one
// 8-byte read-back that picks the radix passes).  *d_depth (device, zeroed by the caller) receives the depth of the tree in
// node pairs: the traversal stack needs one entry per level.  use_ploc: build the hierarchy by parallel locally-ordered clustering
// over the Morton order (better trees, ~40 rounds of 4 small kernels) instead of Karras' one-pass radix-tree construction.
int BuildLbvhOnDevice(const BuildPrim* d_prims, uint32_t n, uint32_t pair_base, uint32_t ref_base, void* d_nodes, uint32_t* d_prim_refs,
                      LbvhScratch* scratch, void* stream, uint64_t* launches, uint32_t* d_depth, bool use_ploc, std::string* err);
void FreeLbvhScratch(LbvhScratch* s);

}  // namespace rt2
