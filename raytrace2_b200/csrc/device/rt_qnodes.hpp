// Compact node pairs for the traversal (implementation: rt_qnodes.cu).
#pragma once
#include <cstdint>
#include <string>

namespace rt2 {

struct NodeGrid {
  float base[3];   // grid origin - extent: a stored coordinate v in [1, 2) means  x = base + v * ext
  float ext[3];
  double inflation;  // mean over the child boxes of (surface area quantised / surface area float) - 1
  bool usable;       // false: a node pair with two empty slots, or non-finite bounds — keep the float nodes
};

// Quantises every node pair of `d_nodes` (4 x float4 per pair, device node format of rt_trace.cuh) onto one 15-bit grid that
// spans the child boxes of the `n_roots` root pairs listed in `roots` (host memory): the trees the traversal will walk.
// `d_qnodes` receives 2 x uint4 per pair.  Synchronous (one small read-back).
int QuantiseNodesOnDevice(const void* d_nodes, uint32_t n_pairs, const uint32_t* roots, uint32_t n_roots, void* d_qnodes, NodeGrid* grid,
                          void* stream, uint64_t* launches, std::string* err);

}  // namespace rt2
