// Binned-SAH BVH2 builder producing 32-byte nodes stored as 64-byte sibling pairs (include/rt2.h: rt2_bvh_node).
//
// This replaces the reference's top-down median split over top-level objects only (src/cpu_raytrace/BVH.cpp:10-31):
// the reference leaves e.g. the 1000-sphere cluster of the book-2 scene as a linear list; here every
// leaf primitive of a space (world or one instance) goes into one SAH tree.  Closest-hit semantics do not depend on
// the tree shape (result = arg-min over leaves of the raw reported t), so only culling quality changes.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "scene_host.hpp"
#include "tuning.hpp"

namespace rt2 {
namespace {

constexpr int kBins = 16;
// Leaf policy: a range of <= MaxLeaf() primitives becomes a leaf when the SAH says splitting does not pay (default 1: on
// the GPU the leaf phase of the while-while walk runs at ~6 of 32 lanes, so one-primitive leaves measured 2-8 % faster
// than leaves of up to 4, profiles/r01_notes.md; longer leaves only appear behind the depth guard), with one
// node-pair visit costed as TravCost() primitive tests.  Overridable only in EXPERIMENTS builds (host/tuning.hpp).
uint32_t MaxLeaf() {
  static const uint32_t v = [] {
    const long x = TuneInt("RT2_BVH_MAX_LEAF", 1);
    return static_cast<uint32_t>(x < 1 ? 1 : (x > 16 ? 16 : x));
  }();
  return v;
}
float TravCost() {
  static const float v = static_cast<float>(TuneFloat("RT2_BVH_TRAV_COST", 2.5));
  return v;
}
constexpr int kMaxDepth = 40;  // beyond this: median splits (<= 24 more levels for 16M prims; the device stack holds 64)

struct Bounds {
  float mn[3], mx[3];
  void Reset() {
    for (int k = 0; k < 3; k++) {
      mn[k] = INFINITY;
      mx[k] = -INFINITY;
    }
  }
  void Grow(const float* a, const float* b) {
    for (int k = 0; k < 3; k++) {
      mn[k] = std::fmin(mn[k], a[k]);
      mx[k] = std::fmax(mx[k], b[k]);
    }
  }
  void Grow(const Bounds& o) { Grow(o.mn, o.mx); }
  float HalfArea() const {
    float dx = mx[0] - mn[0], dy = mx[1] - mn[1], dz = mx[2] - mn[2];
    if (dx < 0 || dy < 0 || dz < 0) return 0.f;
    return dx * dy + dy * dz + dz * dx;
  }
};

struct BuildCtx {
  std::vector<BuildPrim>* prims;
  HostScene* sc;
};

// An empty slot carries NaN bounds: every comparison of the device slab test is then false, so it is never entered.
void SetEmpty(rt2_bvh_node* n) {
  for (int k = 0; k < 3; k++) {
    n->bmin[k] = NAN;
    n->bmax[k] = NAN;
  }
  n->left_first = 0;
  n->count = 0;
}

void SetBox(rt2_bvh_node* n, const Bounds& b) {
  for (int k = 0; k < 3; k++) {
    n->bmin[k] = b.mn[k];
    n->bmax[k] = b.mx[k];
  }
}

Bounds RangeBounds(const std::vector<BuildPrim>& p, size_t begin, size_t end) {
  Bounds b;
  b.Reset();
  for (size_t i = begin; i < end; i++) b.Grow(p[i].bmin, p[i].bmax);
  return b;
}

// Chooses a split of [begin, end); returns the partition point (begin < mid < end).
size_t Split(BuildCtx& ctx, size_t begin, size_t end, const Bounds& bounds, bool* make_leaf) {
  std::vector<BuildPrim>& p = *ctx.prims;
  const size_t n = end - begin;
  *make_leaf = false;
  // An instance reference must sit alone in its leaf: the device traversal switches to the instance's model space
  // when it meets one and does not come back for leaf-mates.
  bool has_instance = false;
  for (size_t i = begin; i < end; i++) has_instance |= (RT2_PRIM_TYPE(p[i].ref) == RT2_PRIM_INSTANCE);
  Bounds cb;
  cb.Reset();
  for (size_t i = begin; i < end; i++) {
    float c[3];
    for (int k = 0; k < 3; k++) c[k] = 0.5f * (p[i].bmin[k] + p[i].bmax[k]);
    cb.Grow(c, c);
  }
  float best_cost = INFINITY;
  int best_axis = -1, best_bin = -1;
  const int axis_first = 0, axis_last = 2;
  for (int axis = axis_first; axis <= axis_last; axis++) {
    float lo = cb.mn[axis], hi = cb.mx[axis];
    if (!(hi > lo)) continue;
    Bounds bin_b[kBins];
    uint32_t bin_n[kBins];
    for (int b = 0; b < kBins; b++) {
      bin_b[b].Reset();
      bin_n[b] = 0;
    }
    float scale = static_cast<float>(kBins) / (hi - lo);
    for (size_t i = begin; i < end; i++) {
      float c = 0.5f * (p[i].bmin[axis] + p[i].bmax[axis]);
      int b = static_cast<int>((c - lo) * scale);
      b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
      bin_b[b].Grow(p[i].bmin, p[i].bmax);
      bin_n[b]++;
    }
    float right_area[kBins];
    uint32_t right_n[kBins];
    Bounds acc;
    acc.Reset();
    uint32_t cnt = 0;
    for (int b = kBins - 1; b > 0; b--) {
      acc.Grow(bin_b[b]);
      cnt += bin_n[b];
      right_area[b] = acc.HalfArea();
      right_n[b] = cnt;
    }
    acc.Reset();
    cnt = 0;
    for (int b = 0; b < kBins - 1; b++) {
      acc.Grow(bin_b[b]);
      cnt += bin_n[b];
      if (cnt == 0 || right_n[b + 1] == 0) continue;
      float cost = acc.HalfArea() * static_cast<float>(cnt) + right_area[b + 1] * static_cast<float>(right_n[b + 1]);
      if (cost < best_cost) {
        best_cost = cost;
        best_axis = axis;
        best_bin = b;
      }
    }
  }
  if (best_axis >= 0) {
    float leaf_cost = bounds.HalfArea() * static_cast<float>(n);
    // traversal cost 1 box-pair test ~ 1.2 primitive tests
    if (!has_instance && n <= MaxLeaf() && leaf_cost <= best_cost + TravCost() * bounds.HalfArea()) {
      *make_leaf = true;
      return begin;
    }
    float lo = cb.mn[best_axis], hi = cb.mx[best_axis];
    float scale = static_cast<float>(kBins) / (hi - lo);
    auto mid_it = std::partition(p.begin() + static_cast<long>(begin), p.begin() + static_cast<long>(end), [&](const BuildPrim& q) {
      float c = 0.5f * (q.bmin[best_axis] + q.bmax[best_axis]);
      int b = static_cast<int>((c - lo) * scale);
      b = b < 0 ? 0 : (b >= kBins ? kBins - 1 : b);
      return b <= best_bin;
    });
    size_t mid = static_cast<size_t>(mid_it - p.begin());
    if (mid > begin && mid < end) return mid;
  }
  // all centroids coincide (or the binned split degenerated)
  if (!has_instance && n <= MaxLeaf()) {
    *make_leaf = true;
    return begin;
  }
  return begin + n / 2;
}

void MakeLeaf(BuildCtx& ctx, rt2_bvh_node* node, size_t begin, size_t end) {
  Bounds b = RangeBounds(*ctx.prims, begin, end);
  SetBox(node, b);
  node->left_first = static_cast<uint32_t>(ctx.sc->prim_refs.size());
  node->count = static_cast<uint32_t>(end - begin);
  for (size_t i = begin; i < end; i++) ctx.sc->prim_refs.push_back((*ctx.prims)[i].ref);
}

void BuildChild(BuildCtx& ctx, uint32_t node_idx, size_t begin, size_t end, int depth) {
  const size_t n = end - begin;
  Bounds b = RangeBounds(*ctx.prims, begin, end);
  bool leaf = (n == 1);
  size_t mid = begin;
  if (!leaf) {
    if (depth >= kMaxDepth) {
      // depth guard: median splits keep the device stack bounded (leaf size stays <= 16, the stack entry's limit)
      if (n <= 2) {
        bool inst = false;
        for (size_t i = begin; i < end; i++) inst |= (RT2_PRIM_TYPE((*ctx.prims)[i].ref) == RT2_PRIM_INSTANCE);
        leaf = !inst;
      }
      mid = begin + n / 2;
    } else {
      mid = Split(ctx, begin, end, b, &leaf);
    }
  }
  if (leaf) {
    MakeLeaf(ctx, &ctx.sc->nodes[node_idx], begin, end);
    return;
  }
  uint32_t child_pair = static_cast<uint32_t>(ctx.sc->nodes.size() / 2);
  ctx.sc->nodes.emplace_back();
  ctx.sc->nodes.emplace_back();
  rt2_bvh_node* node = &ctx.sc->nodes[node_idx];
  SetBox(node, b);
  node->left_first = child_pair;
  node->count = 0;
  BuildChild(ctx, 2 * child_pair, begin, mid, depth + 1);
  BuildChild(ctx, 2 * child_pair + 1, mid, end, depth + 1);
}

}  // namespace

uint32_t BuildBVH(std::vector<BuildPrim>& prims, HostScene* sc) {
  BuildCtx ctx{&prims, sc};
  uint32_t root_pair = static_cast<uint32_t>(sc->nodes.size() / 2);
  sc->nodes.emplace_back();
  sc->nodes.emplace_back();
  SetEmpty(&sc->nodes[2 * root_pair]);
  SetEmpty(&sc->nodes[2 * root_pair + 1]);
  const size_t n = prims.size();
  if (n == 0) return root_pair;
  if (n == 1) {
    MakeLeaf(ctx, &sc->nodes[2 * root_pair], 0, 1);
    return root_pair;
  }
  Bounds b = RangeBounds(prims, 0, n);
  bool leaf = false;
  size_t mid = Split(ctx, 0, n, b, &leaf);
  if (leaf) {
    MakeLeaf(ctx, &sc->nodes[2 * root_pair], 0, n);
    return root_pair;
  }
  BuildChild(ctx, 2 * root_pair, 0, mid, 1);
  BuildChild(ctx, 2 * root_pair + 1, mid, n, 1);
  return root_pair;
}

}  // namespace rt2
