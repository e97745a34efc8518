// CPU helper of tools/tree_quality.py (analysis only, not on any product or test path): builds the radix tree the device LBVH
// builder produces (Karras 2012 == recursive split at the highest differing bit of the sorted Morton keys) and counts the node
// pairs / primitive tests of the walk the extend kernel does (nearest child first, far child pushed, culling against the
// closest hit so far) for a batch of rays.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LEAF 0x80000000u

typedef struct {
  const uint64_t* keys;
  const float* lmin;  // n x 3, sorted order
  const float* lmax;
  float* boxes;       // pairs x 2 x 6  (min xyz, max xyz)
  uint32_t* entries;  // pairs x 2: LEAF | first (one primitive per leaf)  or  child pair
  uint32_t next_pair;
} Build;

static void range_box(const Build* b, int lo, int hi, float* out) {
  for (int k = 0; k < 3; k++) out[k] = INFINITY, out[3 + k] = -INFINITY;
  for (int i = lo; i < hi; i++)
    for (int k = 0; k < 3; k++) {
      if (b->lmin[3 * i + k] < out[k]) out[k] = b->lmin[3 * i + k];
      if (b->lmax[3 * i + k] > out[3 + k]) out[3 + k] = b->lmax[3 * i + k];
    }
}

static int find_split(const Build* b, int lo, int hi) {  // [lo, hi), hi - lo >= 2
  const uint64_t a = b->keys[lo], z = b->keys[hi - 1];
  if (a == z) return (lo + hi) / 2;
  const int bit = 63 - __builtin_clzll(a ^ z);
  const uint64_t mask = 1ull << bit;
  int l = lo, r = hi - 1;  // keys[l] has the bit clear, keys[r] has it set
  while (r - l > 1) {
    const int m = (l + r) / 2;
    if (b->keys[m] & mask) r = m; else l = m;
  }
  return r;
}

// fills slot `slot` (0/1) of pair `pair` with the subtree over [lo, hi); returns the subtree's box in `box`
static void build_rec(Build* b, uint32_t pair, int slot, int lo, int hi, float* box) {
  if (hi - lo == 1) {
    range_box(b, lo, hi, box);
    b->entries[2 * pair + slot] = LEAF | (uint32_t)lo;
  } else {
    const uint32_t child = b->next_pair++;
    const int mid = find_split(b, lo, hi);
    float b0[6], b1[6];
    build_rec(b, child, 0, lo, mid, b0);
    build_rec(b, child, 1, mid, hi, b1);
    for (int k = 0; k < 3; k++) {
      box[k] = b0[k] < b1[k] ? b0[k] : b1[k];
      box[3 + k] = b0[3 + k] > b1[3 + k] ? b0[3 + k] : b1[3 + k];
    }
    b->entries[2 * pair + slot] = child;
  }
  memcpy(b->boxes + 12 * (size_t)pair + 6 * slot, box, 6 * sizeof(float));
}

// n >= 2.  Pair 0 is the root pair.  Returns the number of pairs (n - 1).
uint32_t build_radix_tree(const uint64_t* keys, const float* lmin, const float* lmax, int n, float* boxes, uint32_t* entries) {
  Build b = {keys, lmin, lmax, boxes, entries, 1};
  const int mid = find_split(&b, 0, n);
  float b0[6], b1[6];
  build_rec(&b, 0, 0, 0, mid, b0);
  build_rec(&b, 0, 1, mid, n, b1);
  return b.next_pair;
}

static inline int slab(const float* bx, const float* o, const float* inv, float bound, float* near_out) {
  float tn = 0.0f, tf = INFINITY;
  for (int k = 0; k < 3; k++) {
    float t0 = (bx[k] - o[k]) * inv[k], t1 = (bx[3 + k] - o[k]) * inv[k];
    if (t0 > t1) { float s = t0; t0 = t1; t1 = s; }
    if (t0 > tn) tn = t0;
    if (t1 < tf) tf = t1;
  }
  *near_out = tn;
  const float n0 = tn * 0.999999f;
  return n0 <= tf && n0 <= bound;
}

// leaves: entry = LEAF | first, with `leaf_count[first's leaf]`... kept simple: list leaves are given as (first, count) through
// leaf_first / leaf_count indexed by (entry & ~LEAF) when leaf_count != NULL, else one primitive prim_ids[entry & ~LEAF].
void traverse_count(const float* boxes, const uint32_t* entries, uint32_t root, const uint32_t* prim_ids, const uint32_t* leaf_count,
                    const float* spheres /* n x 4 */, const float* ro, const float* rd, int n_rays, float tmin, float* t_out,
                    int32_t* prim_out, uint32_t* pairs_out, uint32_t* tests_out, uint32_t* depth_out) {
  for (int r = 0; r < n_rays; r++) {
    const float* o = ro + 3 * r;
    const float* d = rd + 3 * r;
    float inv[3];
    for (int k = 0; k < 3; k++) inv[k] = 1.0f / (fabsf(d[k]) > 1e-30f ? d[k] : (d[k] < 0 ? -1e-30f : 1e-30f));
    const double a = (double)d[0] * d[0] + (double)d[1] * d[1] + (double)d[2] * d[2];
    float best = INFINITY;
    int32_t best_prim = -1;
    uint32_t stack[128];
    int sp = 0, max_sp = 0;
    uint32_t cur = root, pairs = 0, tests = 0;
    for (;;) {
      if (!(cur & LEAF)) {
        pairs++;
        float n0, n1;
        const int h0 = slab(boxes + 12 * (size_t)cur, o, inv, best, &n0);
        const int h1 = slab(boxes + 12 * (size_t)cur + 6, o, inv, best, &n1);
        const uint32_t e0 = entries[2 * cur], e1 = entries[2 * cur + 1];
        if (h0 && h1) {
          const int swap = n1 < n0;
          stack[sp++] = swap ? e0 : e1;
          if (sp > max_sp) max_sp = sp;
          cur = swap ? e1 : e0;
          continue;
        }
        if (h0) { cur = e0; continue; }
        if (h1) { cur = e1; continue; }
      } else {
        const uint32_t first = cur & ~LEAF;
        const uint32_t count = leaf_count ? leaf_count[first] : 1u;
        for (uint32_t i = 0; i < count; i++) {
          const uint32_t p = prim_ids[first + i];
          const float* s = spheres + 4 * (size_t)p;
          tests++;
          const double ocx = s[0] - o[0], ocy = s[1] - o[1], ocz = s[2] - o[2];
          const double h = d[0] * ocx + d[1] * ocy + d[2] * ocz;
          const double c = ocx * ocx + ocy * ocy + ocz * ocz - (double)s[3] * s[3];
          const double disc = h * h - a * c;
          if (disc < 0) continue;
          const double sq = sqrt(disc);
          double root_t = (h - sq) / a;
          if (!(tmin < root_t && root_t < best)) {
            root_t = (h + sq) / a;
            if (!(tmin < root_t && root_t < best)) continue;
          }
          best = (float)root_t;
          best_prim = (int32_t)p;
        }
      }
      if (sp == 0) break;
      cur = stack[--sp];
    }
    t_out[r] = best;
    prim_out[r] = best_prim;
    pairs_out[r] = pairs;
    tests_out[r] = tests;
    depth_out[r] = (uint32_t)max_sp;
  }
}

// PLOC (Meister & Bittner 2018), serial restatement: clusters in Morton order, every cluster looks `radius` slots to either side for the
// neighbour with the smallest merged surface area, mutual nearest neighbours merge, the list is compacted, repeat.
// Returns the ROOT pair index (the last pair created); pairs = n - 1.
static inline float merged_area(const float* amin, const float* amax, const float* bmin, const float* bmax) {
  float e[3];
  for (int k = 0; k < 3; k++) e[k] = (amax[k] > bmax[k] ? amax[k] : bmax[k]) - (amin[k] < bmin[k] ? amin[k] : bmin[k]);
  return e[0] * e[1] + e[1] * e[2] + e[2] * e[0];
}

uint32_t build_ploc(const float* lmin, const float* lmax, int n, int radius, float* boxes, uint32_t* entries) {
  float* cmin = malloc(sizeof(float) * 3 * (size_t)n);
  float* cmax = malloc(sizeof(float) * 3 * (size_t)n);
  uint32_t* cent = malloc(sizeof(uint32_t) * (size_t)n);
  int* nn = malloc(sizeof(int) * (size_t)n);
  memcpy(cmin, lmin, sizeof(float) * 3 * (size_t)n);
  memcpy(cmax, lmax, sizeof(float) * 3 * (size_t)n);
  for (int i = 0; i < n; i++) cent[i] = LEAF | (uint32_t)i;
  int m = n;
  uint32_t next_pair = 0;
  while (m > 1) {
    for (int i = 0; i < m; i++) {
      float best = INFINITY;
      int bj = -1;
      const int lo = i - radius < 0 ? 0 : i - radius, hi = i + radius >= m ? m - 1 : i + radius;
      for (int j = lo; j <= hi; j++) {
        if (j == i) continue;
        const float a = merged_area(cmin + 3 * i, cmax + 3 * i, cmin + 3 * j, cmax + 3 * j);
        if (a < best) best = a, bj = j;
      }
      nn[i] = bj;
    }
    int out = 0;
    for (int i = 0; i < m; i++) {
      const int j = nn[i];
      if (nn[j] == i) {
        if (i > j) continue;  // merged into slot j's pair below (j < i handled when we were at j)
        const uint32_t p = next_pair++;
        for (int k = 0; k < 3; k++) {
          boxes[12 * (size_t)p + k] = cmin[3 * i + k], boxes[12 * (size_t)p + 3 + k] = cmax[3 * i + k];
          boxes[12 * (size_t)p + 6 + k] = cmin[3 * j + k], boxes[12 * (size_t)p + 9 + k] = cmax[3 * j + k];
        }
        entries[2 * p] = cent[i], entries[2 * p + 1] = cent[j];
        float umin[3], umax[3];
        for (int k = 0; k < 3; k++) {
          umin[k] = cmin[3 * i + k] < cmin[3 * j + k] ? cmin[3 * i + k] : cmin[3 * j + k];
          umax[k] = cmax[3 * i + k] > cmax[3 * j + k] ? cmax[3 * i + k] : cmax[3 * j + k];
        }
        // (writing slot `out` <= i is safe: slots below i are done; slot j > i is still read-only until we pass it)
        for (int k = 0; k < 3; k++) cmin[3 * out + k] = umin[k], cmax[3 * out + k] = umax[k];
        cent[out] = p;
        // nn[] of later slots refers to OLD indices: keep nn intact, it is only read at old indices >= i
        out++;
      } else {
        if (out != i) {
          for (int k = 0; k < 3; k++) cmin[3 * out + k] = cmin[3 * i + k], cmax[3 * out + k] = cmax[3 * i + k];
          cent[out] = cent[i];
        }
        out++;
      }
    }
    m = out;
  }
  const uint32_t root = cent[0];
  free(cmin); free(cmax); free(cent); free(nn);
  return root;
}
