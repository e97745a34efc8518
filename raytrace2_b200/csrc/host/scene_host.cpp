// Host scene compiler (see scene_host.hpp).  Citations are relative to the reference root.
//
//   JSON  --ParseDocument-->  textures / materials / primitive definitions / node tree   (Serialize.cpp:199-360)
//         --ReplayReferenceBVH--> span-1 ("Q2") flag per top-level node                  (App.cpp:126, BVH.cpp:10-38)
//         --Flatten-->         SoA spheres / quads / instance chains / media + BVH       (new)
//
// The loader keeps the reference's defaults and quirks (int-typed fov default truncates, `texture` / `diffuse_light`
// with an inline albedo append a SolidColor texture, `constant_medium.albedo` appends texture + isotropic material,
// invalid primitives are skipped so later indices shift, absent rotation = identity).  Where the reference would hit
// undefined behaviour or an uncaught exception (out-of-range indices, legacy files) we return an error or adapt, and
// say so in a warning.
#include "scene_host.hpp"

#include "image_in.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <sstream>

#include "json.hpp"

namespace rt2 {
namespace {

// ---------------------------------------------------------------------------------------------------------------
// small deterministic generator for the Perlin tables and the synthetic scene (the reference uses minstd_rand seeded
// from random_device, Math.hpp:9-13 — not reproducible, so any uniform generator is distributionally equivalent)
struct SplitMix64 {
  uint64_t s;
  explicit SplitMix64(uint64_t seed) : s(seed) {}
  uint64_t Next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  float Real() { return static_cast<float>(Next() >> 40) * (1.0f / 16777216.0f); }  // [0,1)
  float Real(float lo, float hi) { return lo + Real() * (hi - lo); }                // Math.hpp:15
  int Int(int lo, int hi) { return static_cast<int>(Real(static_cast<float>(lo), static_cast<float>(hi + 1))); }  // Math.hpp:18
};

// PerlinNoiseGen::Init + GeneratePerm (PerlinNoiseGen.cpp:41-50, 90-103)
void InitPerlin(rt2_perlin* p, int point_count, SplitMix64& rng) {
  std::memset(p, 0, sizeof(*p));
  if (point_count > 256) point_count = 256;  // the hash mask is hard-coded to 255 (PerlinNoiseGen.cpp:83)
  for (int i = 0; i < point_count; i++) {
    V3 v{rng.Real(-1, 1), rng.Real(-1, 1), rng.Real(-1, 1)};
    v = Normalize(v);
    p->vec[i][0] = v.x;
    p->vec[i][1] = v.y;
    p->vec[i][2] = v.z;
  }
  int32_t* perms[3] = {p->perm_x, p->perm_y, p->perm_z};
  for (auto* perm : perms) {
    for (int i = 0; i < point_count; i++) perm[i] = i;
    for (int i = point_count - 1; i > 0; i--) {
      int target = rng.Int(0, i);
      if (target > i) target = i;
      std::swap(perm[i], perm[target]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct PrimDef {
  std::vector<uint32_t> refs;  // sphere: 1, quad: 1, box: 6 quads (MakeBox order, Quad.hpp:43-48)
  RefAABB ref_box;             // the reference's GetAABB() for this primitive
  bool is_medium{false};
  float neg_inv_density{0};
  uint32_t medium_material{0};
};

struct Node {
  int primitive{-1};
  bool has_transform{false};
  M4 model;
  bool has_children{false};
  std::vector<Node> children;
};

struct Builder {
  HostScene* sc;
  std::vector<PrimDef> prims;
  std::vector<Node> top;
  std::string err;

  void Warn(const std::string& m) { sc->warnings.push_back(m); }

  // Quad ctor + SetBoundingBox (Quad.hpp:14-26)
  uint32_t AddQuad(V3 q, V3 u, V3 v, uint32_t material, RefAABB* box) {
    rt2_quad g{};
    V3 n = Cross(u, v);
    V3 normal = Normalize(n);
    float d = Dot(normal, q);
    V3 w = n / Dot(n, n);
    for (int i = 0; i < 3; i++) {
      g.normal[i] = normal[i];
      g.q[i] = q[i];
      g.u[i] = u[i];
      g.v[i] = v[i];
      g.w[i] = w[i];
    }
    g.d = d;
    g.material = material;
    // Axis code for the device's Quad::Hit (rt_trace.cuh quad_hit): 1 + c when normal and w have their only non-zero
    // component on axis c and u, v one non-zero component each (every face of a MakeBox and every Cornell wall).  The
    // reference's arithmetic then multiplies by exact zeros everywhere else, so the terms that survive can be evaluated
    // alone with bit-identical results.
    {
      auto single = [](const V3& a, int* axis) {
        int n = 0;
        for (int i = 0; i < 3; i++)
          if (a[i] != 0.0f) *axis = i, n++;
        return n == 1;
      };
      int cn = -1, cw = -1, cu = -1, cv = -1;
      uint32_t code = 0;
      if (single(normal, &cn) && single(w, &cw) && cn == cw && single(u, &cu) && single(v, &cv) && cu != cn && cv != cn && cu != cv)
        code = 1u + static_cast<uint32_t>(cn);
      std::memcpy(&g.pad0, &code, sizeof(code));
    }
    *box = RefAABB{RefAABB{q, q + u + v}, RefAABB{q + u, q + v}};
    sc->quads.push_back(g);
    return (RT2_PRIM_QUAD << 28) | static_cast<uint32_t>(sc->quads.size() - 1);
  }
  // Sphere moving ctor (Sphere.hpp:21-29); the loader always uses it (Serialize.cpp:311-314)
  uint32_t AddSphere(V3 c, V3 disp, float radius, uint32_t material, RefAABB* box) {
    rt2_sphere s{};
    for (int i = 0; i < 3; i++) {
      s.center0[i] = c[i];
      s.displacement[i] = disp[i];
    }
    s.radius = radius;
    s.material = material;
    V3 c0 = c + disp * 0.0f, c1 = c + disp * 1.0f;  // Ray::At(0), Ray::At(1)
    V3 r{radius, radius, radius};
    *box = RefAABB{RefAABB{c0 - r, c0 + r}, RefAABB{c1 - r, c1 + r}};
    sc->spheres.push_back(s);
    return (RT2_PRIM_SPHERE << 28) | static_cast<uint32_t>(sc->spheres.size() - 1);
  }
  // MakeBox (Quad.hpp:34-50): HittableList of 6 quads; list AABB = running padded union (HittableList.hpp:13-16)
  void AddBox(V3 a, V3 b, uint32_t material, PrimDef* def) {
    V3 mn{std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)};
    V3 mx{std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)};
    V3 dx{mx.x - mn.x, 0, 0}, dy{0, mx.y - mn.y, 0}, dz{0, 0, mx.z - mn.z};
    struct Q { V3 q, u, v; };
    const Q faces[6] = {{V3{mn.x, mn.y, mx.z}, dx, dy},  {V3{mx.x, mn.y, mx.z}, -dz, dy}, {V3{mx.x, mn.y, mn.z}, -dx, dy},
                        {V3{mn.x, mn.y, mn.z}, dz, dy},  {V3{mn.x, mx.y, mx.z}, dx, -dz}, {V3{mn.x, mn.y, mn.z}, dx, dz}};
    RefAABB list_box;
    for (const Q& f : faces) {
      RefAABB qb;
      def->refs.push_back(AddQuad(f.q, f.u, f.v, material, &qb));
      list_box = RefAABB{list_box, qb};
    }
    def->ref_box = list_box;
  }
};

bool ReadV3(const json::Value& obj, const char* key, V3 def, V3* out) {
  float a[3];
  if (obj.GetFloatArray<3>(key, a)) {
    *out = V3{a[0], a[1], a[2]};
    return true;
  }
  *out = def;
  return false;
}

// LoadCamera (Serialize.cpp:32-40)
CameraParams ParseCamera(const json::Value& obj) {
  CameraParams c;
  c.vfov = static_cast<float>(obj.GetInt("fov", 90));  // int default: a fractional fov is truncated
  ReadV3(obj, "center", V3{0, 0, 1}, &c.center);
  ReadV3(obj, "look_at", V3{0, 0, 0}, &c.look_at);
  c.defocus_angle = obj.GetFloat("defocus_angle", 0.0f);
  c.focus_dist = obj.GetFloat("focus_distance", 1.f);
  return c;
}

bool ReadFile(const std::string& path, std::string* out) {
  std::ifstream f(path, std::ios::binary);
  if (!f.is_open()) return false;
  std::stringstream ss;
  ss << f.rdbuf();
  *out = ss.str();
  return true;
}

int LoadCameraFile(const std::string& path, CameraParams* cam, std::string* err) {
  std::string text;
  if (!ReadFile(path, &text)) {
    *err = "Failed to open json file: " + path;
    return RT2_ERR_IO;
  }
  json::Value v;
  std::string perr;
  if (!json::Parser(text).Parse(v, &perr) || !v.IsObject()) {
    *err = "camera file is not a JSON object: " + path;
    return RT2_ERR_PARSE;
  }
  *cam = ParseCamera(v);
  return RT2_OK;
}

// ParseTransform (Serialize.cpp:106-132): M = T * R * S; rotation = [angle_deg, ax, ay, az], axis not normalised;
// an absent rotation leaves glm::quat uninitialised in the reference (:114) — defined here as identity.
bool ParseTransform(const json::Value& node, M4* out) {
  const json::Value* t = node.Find("transform");
  if (!t || !t->IsObject()) return false;
  V3 translation{0, 0, 0};
  ReadV3(*t, "translation", V3{0, 0, 0}, &translation);
  Quat rot;  // identity
  float aa[4];
  if (t->Contains("rotation")) {
    if (!t->GetFloatArray<4>("rotation", aa)) {
      aa[0] = 0, aa[1] = 0, aa[2] = 1, aa[3] = 0;  // the reference's default {0,0,1,0}
    }
    rot = AngleAxis(Radians(aa[0]), V3{aa[1], aa[2], aa[3]});
  }
  V3 scale{1, 1, 1};
  ReadV3(*t, "scale", V3{1, 1, 1}, &scale);
  *out = Mul(Mul(Translate(M4::Identity(), translation), ToMat4(rot)), Scale(M4::Identity(), scale));
  return true;
}

// ParseNode (Serialize.cpp:161-197)
int ParseNode(const json::Value& j, size_t n_prims, Node* out, std::string* err) {
  if (j.Contains("primitive")) {
    int idx = j.GetInt("primitive", -1);
    if (idx < 0 || idx >= static_cast<int>(n_prims)) {
      *err = "primitive out of range of primitives";  // the reference prints this and then indexes out of range
      return RT2_ERR_PARSE;
    }
    out->primitive = idx;
  }
  out->has_transform = ParseTransform(j, &out->model);
  if (const json::Value* ch = j.Find("children")) {
    if (!ch->IsArray()) {
      // reference: prints "children entry must be an array" and carries on with the bare primitive
    } else {
      out->has_children = true;
      for (size_t i = 0; i < ch->Size(); i++) {
        Node c;
        int rc = ParseNode(ch->At(i), n_prims, &c, err);
        if (rc != RT2_OK) return rc;
        out->children.push_back(std::move(c));
      }
    }
  }
  if (out->primitive < 0 && !out->has_children) {
    *err = "error parsing node";  // reference: prints, then dereferences a null Hittable
    return RT2_ERR_PARSE;
  }
  return RT2_OK;
}

// The reference's GetAABB() of a parsed node (HittableList.hpp:13-16, Transform.cpp:36-64).
RefAABB NodeRefBox(const Node& n, const std::vector<PrimDef>& prims) {
  RefAABB box;
  bool single = (n.primitive >= 0 && !n.has_children);
  if (single) {
    box = prims[n.primitive].ref_box;
  } else {
    if (n.primitive >= 0) box = RefAABB{box, prims[n.primitive].ref_box};
    for (const Node& c : n.children) box = RefAABB{box, NodeRefBox(c, prims)};
  }
  if (n.has_transform) {
    V3 mn = box.Min(), mx = box.Max();
    V3 corners[8] = {V3{mn.x, mn.y, mn.z}, V3{mx.x, mn.y, mn.z}, V3{mn.x, mx.y, mn.z}, V3{mx.x, mx.y, mn.z},
                     V3{mn.x, mn.y, mx.z}, V3{mx.x, mn.y, mx.z}, V3{mn.x, mx.y, mx.z}, V3{mx.x, mx.y, mx.z}};
    V3 nmin{kInfinity, kInfinity, kInfinity}, nmax{-kInfinity, -kInfinity, -kInfinity};
    for (const V3& c : corners) {
      V3 t = MulPoint(n.model, c);
      for (int k = 0; k < 3; k++) {
        nmin[k] = std::fmin(nmin[k], t[k]);
        nmax[k] = std::fmax(nmax[k], t[k]);
      }
    }
    box = RefAABB{nmin, nmax};
  }
  return box;
}

// Replays BVHNode's constructor (BVH.cpp:10-31) on the top-level list to find the objects that end up in span-1
// leaves (left_ == right_ == object => Hit() is called twice, BVH.cpp:50-55).  std::sort is libstdc++'s introsort in
// both builds; its permutation depends only on the comparator outcomes, which are reproduced bit-for-bit.
struct ReplayItem {
  RefAABB box;
  uint32_t idx;
};
void ReplayReferenceBVH(std::vector<ReplayItem>& objs, size_t start, size_t end, std::vector<uint8_t>& span1) {
  RefAABB box;
  for (size_t i = start; i < end; i++) box = RefAABB{box, objs[i].box};
  size_t span = end - start;
  if (span == 1) {
    span1[objs[start].idx] = 1;
  } else if (span == 2) {
    // one object per side: no duplicate
  } else {
    int axis = box.LongestAxis();
    std::sort(objs.begin() + static_cast<long>(start), objs.begin() + static_cast<long>(end),
              [axis](const ReplayItem& a, const ReplayItem& b) { return a.box.Axis(axis).min < b.box.Axis(axis).min; });
    size_t mid = start + span / 2;
    ReplayReferenceBVH(objs, start, mid, span1);
    ReplayReferenceBVH(objs, mid, end, span1);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Flattening
struct Box3 {
  float mn[3]{kInfinity, kInfinity, kInfinity};
  float mx[3]{-kInfinity, -kInfinity, -kInfinity};
  void Grow(const float* p) {
    for (int k = 0; k < 3; k++) {
      mn[k] = std::fmin(mn[k], p[k]);
      mx[k] = std::fmax(mx[k], p[k]);
    }
  }
  void Grow(const Box3& b) {
    Grow(b.mn);
    Grow(b.mx);
  }
};

// Conservative float bounds of a primitive in its own space, padded by a few ulps so the traversal's slab test never
// culls a hit the reference's exact arithmetic would report.
Box3 PrimBounds(const HostScene& sc, uint32_t ref) {
  Box3 b;
  uint32_t i = RT2_PRIM_INDEX(ref);
  if (RT2_PRIM_TYPE(ref) == RT2_PRIM_SPHERE) {
    const rt2_sphere& s = sc.spheres[i];
    float r = std::fabs(s.radius);
    for (int e = 0; e < 2; e++) {
      for (int k = 0; k < 3; k++) {
        float c = s.center0[k] + s.displacement[k] * static_cast<float>(e);
        float lo = c - r, hi = c + r;
        b.mn[k] = std::fmin(b.mn[k], lo);
        b.mx[k] = std::fmax(b.mx[k], hi);
      }
    }
  } else {
    const rt2_quad& q = sc.quads[i];
    for (int cu = 0; cu < 2; cu++)
      for (int cv = 0; cv < 2; cv++) {
        float p[3];
        for (int k = 0; k < 3; k++) p[k] = q.q[k] + (cu ? q.u[k] : 0.f) + (cv ? q.v[k] : 0.f);
        b.Grow(p);
      }
  }
  for (int k = 0; k < 3; k++) {
    float ext = std::fmax(std::fabs(b.mn[k]), std::fabs(b.mx[k]));
    float pad = ext * 4e-7f + 5e-5f;  // >= the reference's 1e-4 minimum thickness (AABB.hpp:58-64)
    b.mn[k] -= pad;
    b.mx[k] += pad;
  }
  return b;
}

constexpr size_t kMaxUnifiedLeaves = 1u << 22;

struct Flattener {
  HostScene* sc;
  const std::vector<PrimDef>* prims;
  std::vector<BuildPrim> tlas;
  std::vector<BuildPrim> unified_leaves;  // the instanced primitives as world-space leaves (see HostScene::inst_leaves)

  // Conservative world-space box of primitive `ref` under the instance chain (outermost level first).
  Box3 InstancedPrimWorldBounds(uint32_t ref, const std::vector<M4>& chain, double sigma_max) const {
    Box3 b;
    auto to_world = [&](V3 p) {
      for (size_t l = chain.size(); l-- > 0;) p = MulPoint(chain[l], p);
      return p;
    };
    const uint32_t i = RT2_PRIM_INDEX(ref);
    if (RT2_PRIM_TYPE(ref) == RT2_PRIM_SPHERE) {
      // a sphere maps to an ellipsoid inside the ball of radius r * sigma_max(chain) around the mapped centre
      const rt2_sphere& s = sc->spheres[i];
      const float r = static_cast<float>(std::fabs(s.radius) * sigma_max * 1.000001);
      for (int e = 0; e < 2; e++) {
        const V3 c = to_world(V3{s.center0[0] + s.displacement[0] * static_cast<float>(e), s.center0[1] + s.displacement[1] * static_cast<float>(e),
                                 s.center0[2] + s.displacement[2] * static_cast<float>(e)});
        const float lo[3] = {c.x - r, c.y - r, c.z - r}, hi[3] = {c.x + r, c.y + r, c.z + r};
        b.Grow(lo);
        b.Grow(hi);
      }
    } else {
      const rt2_quad& q = sc->quads[i];
      for (int cu = 0; cu < 2; cu++)
        for (int cv = 0; cv < 2; cv++) {
          const V3 w = to_world(V3{q.q[0] + (cu ? q.u[0] : 0.f) + (cv ? q.v[0] : 0.f), q.q[1] + (cu ? q.u[1] : 0.f) + (cv ? q.v[1] : 0.f),
                                   q.q[2] + (cu ? q.u[2] : 0.f) + (cv ? q.v[2] : 0.f)});
          const float p[3] = {w.x, w.y, w.z};
          b.Grow(p);
        }
    }
    for (int k = 0; k < 3; k++) {
      const float ext = std::fmax(std::fabs(b.mn[k]), std::fabs(b.mx[k]));
      const float pad = ext * 2e-6f + 1e-4f;
      b.mn[k] -= pad;
      b.mx[k] += pad;
    }
    return b;
  }

  static BuildPrim MakeBuildPrim(const Box3& b, uint32_t ref) {
    BuildPrim p;
    for (int k = 0; k < 3; k++) {
      p.bmin[k] = b.mn[k];
      p.bmax[k] = b.mx[k];
    }
    p.ref = ref;
    return p;
  }

  void AddPrimitive(int prim_idx, const std::vector<M4>& chain, std::vector<BuildPrim>* target, uint32_t top_idx) {
    const PrimDef& def = (*prims)[prim_idx];
    if (def.is_medium) {
      rt2_medium m{};
      m.neg_inv_density = def.neg_inv_density;
      m.material = def.medium_material;
      m.boundary_first = static_cast<uint32_t>(sc->prim_refs.size());
      m.boundary_count = static_cast<uint32_t>(def.refs.size());
      for (uint32_t r : def.refs) sc->prim_refs.push_back(r);
      m.chain_first = static_cast<uint32_t>(sc->xforms.size());
      m.chain_len = static_cast<uint32_t>(chain.size());
      PushChain(chain);
      m.sample_twice = sc->span1_flags[top_idx];
      m.top_level_node = top_idx;
      sc->media.push_back(m);
      // conservative box of the boundary (device-side early out of ConstantMedium::Hit, rt_trace.cuh finish_hit)
      Box3 mb = PrimBounds(*sc, def.refs[0]);
      for (uint32_t r : def.refs) {
        const Box3 pb = PrimBounds(*sc, r);
        for (int k = 0; k < 3; k++) {
          mb.mn[k] = std::min(mb.mn[k], pb.mn[k]);
          mb.mx[k] = std::max(mb.mx[k], pb.mx[k]);
        }
      }
      for (int k = 0; k < 3; k++) sc->media_bounds.push_back(mb.mn[k]);
      sc->media_bounds.push_back(0.f);
      for (int k = 0; k < 3; k++) sc->media_bounds.push_back(mb.mx[k]);
      sc->media_bounds.push_back(0.f);
      return;
    }
    for (uint32_t r : def.refs) target->push_back(MakeBuildPrim(PrimBounds(*sc, r), r));
  }

  void PushChain(const std::vector<M4>& chain) {
    for (const M4& model : chain) {
      M4 inv = Inverse(model);  // TransformedHittable::Init, Transform.cpp:37
      rt2_xform x{};
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 4; c++) {
          x.inv[r][c] = inv[c][r];
          x.model[r][c] = model[c][r];
        }
      sc->xforms.push_back(x);
      // smallest singular value of the inverse 3x3 (for conservative TLAS culling under Q1)
      double a[3][3];
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) {
          double s = 0;
          for (int k = 0; k < 3; k++) s += static_cast<double>(inv[r][k]) * inv[c][k];  // (inv^T inv)[r][c] with inv[c][r] storage
          a[r][c] = s;
        }
      level_sigma_.push_back(SmallestEig3(a));
    }
  }

  // smallest eigenvalue of a symmetric 3x3 via cyclic Jacobi; returns sqrt (= singular value)
  static double SmallestEig3(double a[3][3]) {
    for (int sweep = 0; sweep < 32; sweep++) {
      double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
      if (off < 1e-15) break;
      for (int p = 0; p < 2; p++)
        for (int q = p + 1; q < 3; q++) {
          if (std::fabs(a[p][q]) < 1e-300) continue;
          double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
          double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
          double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
          for (int k = 0; k < 3; k++) {
            double akp = a[k][p], akq = a[k][q];
            a[k][p] = c * akp - s * akq;
            a[k][q] = s * akp + c * akq;
          }
          for (int k = 0; k < 3; k++) {
            double apk = a[p][k], aqk = a[q][k];
            a[p][k] = c * apk - s * aqk;
            a[q][k] = s * apk + c * aqk;
          }
        }
    }
    double e = std::fmin(a[0][0], std::fmin(a[1][1], a[2][2]));
    return std::sqrt(std::fmax(e, 0.0));
  }

  void FlattenNode(const Node& n, std::vector<M4> chain, std::vector<BuildPrim>* target, uint32_t top_idx) {
    if (!n.has_transform) {
      if (n.primitive >= 0) AddPrimitive(n.primitive, chain, target, top_idx);
      for (const Node& c : n.children) FlattenNode(c, chain, target, top_idx);
      return;
    }
    chain.push_back(n.model);
    std::vector<BuildPrim> own;
    if (n.primitive >= 0) AddPrimitive(n.primitive, chain, &own, top_idx);
    for (const Node& c : n.children) FlattenNode(c, chain, &own, top_idx);
    if (own.empty()) return;  // only media (or nested instances) below this transform
    // model-space bounds of everything this instance owns
    Box3 local;
    for (const BuildPrim& p : own) {
      local.Grow(p.bmin);
      local.Grow(p.bmax);
    }
    rt2_instance inst{};
    inst.chain_first = static_cast<uint32_t>(sc->xforms.size());
    inst.chain_len = static_cast<uint32_t>(chain.size());
    size_t sigma_first = level_sigma_.size();
    PushChain(chain);
    inst.top_level_node = top_idx;
    inst.blas_root = BuildBVH(own, sc);
    sc->instances.push_back(inst);
    sc->tree_prims.resize(sc->instances.size() + 1);
    sc->tree_prims[sc->instances.size()] = own;
    double sigma = 1.0;
    for (size_t i = sigma_first; i < level_sigma_.size(); i++) sigma *= level_sigma_[i];
    sc->min_inv_scale = std::fmin(sc->min_inv_scale, static_cast<float>(sigma * 0.9999));
    // unified world tree: this instance's primitives as world-space leaves (sigma = sigma_min of the inverse chain, so
    // 1 / sigma bounds the largest stretch of the model matrix)
    if (sigma > 0.0) {
      const uint32_t inst_idx = static_cast<uint32_t>(sc->instances.size() - 1);
      for (const BuildPrim& p : own) {
        if (RT2_PRIM_TYPE(p.ref) != RT2_PRIM_SPHERE && RT2_PRIM_TYPE(p.ref) != RT2_PRIM_QUAD) continue;
        const uint32_t k = static_cast<uint32_t>(sc->inst_leaves.size() / 2);
        sc->inst_leaves.push_back(p.ref);
        sc->inst_leaves.push_back(inst_idx);
        unified_leaves.push_back(MakeBuildPrim(InstancedPrimWorldBounds(p.ref, chain, 1.0 / sigma), (RT2_PRIM_INSTANCE << 28) | k));
      }
    } else {
      unified_ok = false;  // singular transform: no finite world bound
    }
    // world bounds: the 8 corners through the chain, innermost level first (p_world = M_0 * (M_1 * ... p))
    Box3 world;
    for (int ci = 0; ci < 8; ci++) {
      V3 p{(ci & 1) ? local.mx[0] : local.mn[0], (ci & 2) ? local.mx[1] : local.mn[1], (ci & 4) ? local.mx[2] : local.mn[2]};
      for (size_t l = chain.size(); l-- > 0;) p = MulPoint(chain[l], p);
      float pp[3] = {p.x, p.y, p.z};
      world.Grow(pp);
    }
    for (int k = 0; k < 3; k++) {
      float ext = std::fmax(std::fabs(world.mn[k]), std::fabs(world.mx[k]));
      float pad = ext * 2e-6f + 1e-4f;
      world.mn[k] -= pad;
      world.mx[k] += pad;
    }
    tlas.push_back(MakeBuildPrim(world, (RT2_PRIM_INSTANCE << 28) | static_cast<uint32_t>(sc->instances.size() - 1)));
    for (int k = 0; k < 3; k++) sc->inst_bounds.push_back(world.mn[k]);
    sc->inst_bounds.push_back(0.f);
    for (int k = 0; k < 3; k++) sc->inst_bounds.push_back(world.mx[k]);
    sc->inst_bounds.push_back(0.f);
  }

  std::vector<double> level_sigma_;
  bool unified_ok{true};
};

int Compile(Builder& b, std::string* err) {
  HostScene* sc = b.sc;
  sc->n_top_level = static_cast<uint32_t>(b.top.size());
  // Q2 flags from the reference-order BVH
  sc->span1_flags.assign(b.top.size(), 0);
  std::vector<ReplayItem> ref_order(b.top.size());  // the top-level objects in the reference BVH's leaf order
  if (!b.top.empty()) {
    for (size_t i = 0; i < b.top.size(); i++) ref_order[i] = ReplayItem{NodeRefBox(b.top[i], b.prims), static_cast<uint32_t>(i)};
    ReplayReferenceBVH(ref_order, 0, ref_order.size(), sc->span1_flags);
  }
  // Tie ranks (device quad_wins_tie): position of every quad in the order in which the reference's BVH visits the leaves —
  // its top-level objects left to right (BVH.cpp:50-55: always left, then right), each object's own list in list order
  // (own primitive, then children; HittableList.cpp:8-22).  Stored as integer bits in rt2_quad.pad1.
  {
    uint32_t next_rank = 1;
    std::vector<uint8_t> ranked(sc->quads.size(), 0);
    auto rank_node = [&](const Node& n, auto&& self) -> void {
      if (n.primitive >= 0) {
        for (uint32_t r : b.prims[static_cast<size_t>(n.primitive)].refs) {
          if (RT2_PRIM_TYPE(r) != RT2_PRIM_QUAD || ranked[RT2_PRIM_INDEX(r)]) continue;
          ranked[RT2_PRIM_INDEX(r)] = 1;
          const uint32_t v = next_rank++;
          std::memcpy(&sc->quads[RT2_PRIM_INDEX(r)].pad1, &v, sizeof(v));
        }
      }
      for (const Node& c : n.children) self(c, self);
    };
    for (const ReplayItem& it : ref_order) rank_node(b.top[it.idx], rank_node);
  }
  Flattener fl{sc, &b.prims, {}, {}, {}};
  sc->min_inv_scale = 1.0f;
  for (size_t i = 0; i < b.top.size(); i++) fl.FlattenNode(b.top[i], {}, &fl.tlas, static_cast<uint32_t>(i));
  sc->tlas_root = BuildBVH(fl.tlas, sc);
  sc->tree_prims.resize(sc->instances.size() + 1);
  sc->tree_prims[0] = fl.tlas;
  // instance split: a second world tree without the instance leaves (RT2_MAX_HOISTED_INSTANCES)
  sc->has_world_tlas = false;
  if (!sc->instances.empty() && sc->instances.size() <= RT2_MAX_HOISTED_INSTANCES) {
    std::vector<BuildPrim> surfaces;
    for (const BuildPrim& p : fl.tlas)
      if (RT2_PRIM_TYPE(p.ref) != RT2_PRIM_INSTANCE) surfaces.push_back(p);
    sc->tlas_world_root = BuildBVH(surfaces, sc);
    sc->has_world_tlas = true;
  }
  // unified world tree: the world surfaces + every instanced primitive as a world-space leaf
  sc->has_unified_tlas = false;
  if (!sc->instances.empty() && fl.unified_ok && !fl.unified_leaves.empty() && fl.unified_leaves.size() <= kMaxUnifiedLeaves) {
    std::vector<BuildPrim> all;
    for (const BuildPrim& p : fl.tlas)
      if (RT2_PRIM_TYPE(p.ref) != RT2_PRIM_INSTANCE) all.push_back(p);
    all.insert(all.end(), fl.unified_leaves.begin(), fl.unified_leaves.end());
    sc->unified_prims = all;
    sc->tlas_unified_root = BuildBVH(all, sc);
    sc->has_unified_tlas = true;
  }
  sc->UpdateCamera();
  (void)err;
  return RT2_OK;
}

// ---------------------------------------------------------------------------------------------------------------
int ParseDocument(const json::Value& obj, const std::string& data_dir, uint64_t perlin_seed, bool is_final_render,
                  HostScene* sc, std::string* err) {
  Builder b;
  b.sc = sc;
  SplitMix64 rng(perlin_seed ^ 0x5EEDF00Dull);
  if (!obj.IsObject()) {
    *err = "scene file is not a JSON object";
    return RT2_ERR_PARSE;
  }
  // background (Serialize.cpp:204)
  V3 bg;
  ReadV3(obj, "background_color", V3{1, 1, 1}, &bg);
  for (int k = 0; k < 3; k++) sc->background[k] = bg[k];

  // camera (Serialize.cpp:203-211)
  const json::Value* cam = obj.Find("camera");
  bool cam_is_object = cam && cam->IsObject();
  if (cam_is_object) {
    sc->cam = ParseCamera(*cam);
  } else if (cam && cam->IsString()) {
    int rc = LoadCameraFile(data_dir + "/" + cam->string + ".json", &sc->cam, err);
    if (rc != RT2_OK) return rc;
  } else {
    // HEAD throws here (type_error.302).  Legacy adapter: book-1 final renders use data/cam1.json,
    // everything else the loader's own camera defaults.
    CameraParams def = ParseCamera(json::Value{});
    if (is_final_render) {
      std::string cerr;
      if (LoadCameraFile(data_dir + "/cam1.json", &def, &cerr) != RT2_OK) {
        def.center = V3{13, 2, 3};
        def.look_at = V3{0, 0, 0};
        def.vfov = 20;
        def.defocus_angle = 0.6f;
        def.focus_dist = 10;
      }
      b.Warn("legacy scene without camera: using the book-1 final camera (data/cam1.json)");
    } else {
      b.Warn("legacy scene without camera: using loader defaults");
    }
    sc->cam = def;
  }

  // textures (Serialize.cpp:216-242)
  if (const json::Value* texs = obj.Find("textures"); texs && texs->IsArray()) {
    for (size_t i = 0; i < texs->Size(); i++) {
      const json::Value& jt = texs->At(i);
      std::string type = jt.GetString("type", "");
      rt2_texture t{};
      V3 albedo;
      if (type == "solid_color") {
        t.type = RT2_TEX_SOLID;
        ReadV3(jt, "albedo", V3{1, 1, 1}, &albedo);
      } else if (type == "checker") {
        t.type = RT2_TEX_CHECKER;
        t.scale = 1.f / jt.GetFloat("scale", 1.0f);  // Checker ctor stores inv_scale (Texture.hpp:20-21)
        t.even_tex_idx = jt.GetUint("even_tex_idx", 0u);
        t.odd_tex_idx = jt.GetUint("odd_tex_idx", 0u);
      } else if (type == "noise") {
        t.type = RT2_TEX_NOISE;
        ReadV3(jt, "albedo", V3{1, 1, 1}, &albedo);
        t.scale = jt.GetFloat("scale", 1.0f);
        t.noise_type = static_cast<uint32_t>(jt.GetInt("noise_type", 1));
        t.perlin_idx = static_cast<uint32_t>(sc->perlin.size());
        rt2_perlin p;
        InitPerlin(&p, jt.GetInt("point_count", 256), rng);
        sc->perlin.push_back(p);
      } else if (type == "image") {
        // Schema extension: {"type": "image", "path": "<file relative to the data directory>"}.  Bytes are taken
        // as sRGB-ish and linearised with the 2.2 power law stb_image's float loader applies (the book the reference follows
        // loads its textures that way); a missing / undecodable file becomes a 1x1 cyan image, the book's debugging aid.
        t.type = RT2_TEX_IMAGE;
        albedo = V3{1, 1, 1};
        std::string rel = jt.GetString("path", "");
        std::string full = (!rel.empty() && rel[0] == '/') ? rel : data_dir + "/" + rel;
        int iw = 0, ih = 0;
        std::vector<uint8_t> rgb;
        std::string ierr;
        rt2_image img{};
        img.texel_offset = static_cast<uint32_t>(sc->image_texels.size() / 4);
        if (rel.empty() || !DecodeImageFile(full, &iw, &ih, &rgb, &ierr)) {
          b.Warn("image texture: " + (rel.empty() ? std::string("no path") : ierr) + " (using solid cyan)");
          iw = ih = 1;
          sc->image_texels.insert(sc->image_texels.end(), {0.f, 1.f, 1.f, 1.f});
        } else {
          float lut[256];
          for (int k = 0; k < 256; k++) lut[k] = std::pow(static_cast<float>(k) / 255.0f, 2.2f);
          sc->image_texels.reserve(sc->image_texels.size() + rgb.size() / 3 * 4);
          for (size_t k = 0; k + 2 < rgb.size(); k += 3) {
            sc->image_texels.push_back(lut[rgb[k]]);
            sc->image_texels.push_back(lut[rgb[k + 1]]);
            sc->image_texels.push_back(lut[rgb[k + 2]]);
            sc->image_texels.push_back(1.0f);
          }
        }
        img.width = static_cast<uint32_t>(iw);
        img.height = static_cast<uint32_t>(ih);
        t.image_idx = static_cast<uint32_t>(sc->images.size());
        sc->images.push_back(img);
      } else {
        // reference: prints "Invalid texture type" and still appends a default-constructed variant (SolidColor)
        b.Warn("Invalid texture type: " + type + " (kept as a black solid colour)");
        t.type = RT2_TEX_SOLID;
        albedo = V3{0, 0, 0};
      }
      for (int k = 0; k < 3; k++) t.albedo[k] = albedo[k];
      sc->textures.push_back(t);
    }
  }

  auto append_solid = [&](V3 albedo) -> uint32_t {
    rt2_texture t{};
    t.type = RT2_TEX_SOLID;
    for (int k = 0; k < 3; k++) t.albedo[k] = albedo[k];
    sc->textures.push_back(t);
    return static_cast<uint32_t>(sc->textures.size() - 1);
  };

  // materials (Serialize.cpp:244-285)
  const json::Value* mats = obj.Find("materials");
  if (!mats || !mats->IsArray()) {
    *err = "scene has no materials array";
    return RT2_ERR_PARSE;
  }
  for (size_t i = 0; i < mats->Size(); i++) {
    const json::Value& jm = mats->At(i);
    std::string type = jm.GetString("type", "");
    rt2_material m{};
    V3 albedo{0, 0, 0};
    if (type.empty()) {
      if (!obj.Contains("scene") && jm.Contains("tex_idx")) {
        // legacy data/final_render_checker.json: {"id":0,"tex_idx":0} — treated as a `texture` material
        b.Warn("legacy material without type: treated as 'texture'");
        type = "texture";
      } else {
        *err = "material type field empty";  // reference: LoadScene returns nullopt -> exit(1)
        return RT2_ERR_PARSE;
      }
    }
    if (type == "lambertian") {
      m.type = RT2_MAT_LAMBERTIAN;
      ReadV3(jm, "albedo", V3{1, 1, 1}, &albedo);
    } else if (type == "dielectric") {
      m.type = RT2_MAT_DIELECTRIC;
      m.refraction_index = jm.GetFloat("refraction_index", 1.0f);
    } else if (type == "metal") {
      m.type = RT2_MAT_METAL;
      ReadV3(jm, "albedo", V3{1, 1, 1}, &albedo);
      m.fuzz = jm.GetFloat("fuzz", 0.0f);
    } else if (type == "texture" || type == "diffuse_light") {
      m.type = (type == "texture") ? RT2_MAT_TEXTURE : RT2_MAT_DIFFUSE_LIGHT;
      if (jm.Contains("tex_idx")) {
        m.tex_idx = jm.GetUint("tex_idx", 0u);
      } else if (jm.Contains("albedo")) {
        V3 a;
        ReadV3(jm, "albedo", V3{1, 1, 1}, &a);
        m.tex_idx = append_solid(a);
      } else {
        // reference: prints an error and keeps the default-constructed variant (a zero MaterialMetal)
        b.Warn("invalid " + type + ", must contain tex_idx or albedo (kept as black metal)");
        m.type = RT2_MAT_METAL;
      }
    } else {
      b.Warn("Invalid material type '" + type + "' (kept as black metal, like the reference's default variant)");
      m.type = RT2_MAT_METAL;
    }
    for (int k = 0; k < 3; k++) m.albedo[k] = albedo[k];
    sc->materials.push_back(m);
  }

  // primitives
  const json::Value* jprims = obj.Find("primitives");
  bool legacy = jprims && jprims->IsObject();
  auto finish_medium = [&](const json::Value& jp, PrimDef* def) -> bool {
    // Serialize.cpp:320-340
    const json::Value* cm = jp.Find("constant_medium");
    if (!cm) return true;
    uint32_t material_idx;
    if (cm->Contains("albedo")) {
      V3 a;
      ReadV3(*cm, "albedo", V3{0, 0, 0}, &a);
      rt2_material iso{};
      iso.type = RT2_MAT_ISOTROPIC;
      iso.tex_idx = append_solid(a);
      material_idx = static_cast<uint32_t>(sc->materials.size());
      sc->materials.push_back(iso);
    } else if (cm->Contains("material")) {
      material_idx = cm->GetUint("material", 0u);
    } else {
      b.Warn("constant_medium must contain 'albedo' or 'material' (primitive skipped)");
      return false;
    }
    float density = static_cast<float>(cm->GetDouble("density", 0.01));
    def->is_medium = true;
    def->neg_inv_density = static_cast<float>(-1.0 / static_cast<double>(density));  // ConstantMedium.cpp:12
    def->medium_material = material_idx;
    return true;
  };

  if (!legacy && jprims && jprims->IsArray()) {
    // Serialize.cpp:287-342
    for (size_t i = 0; i < jprims->Size(); i++) {
      const json::Value& jp = jprims->At(i);
      std::string type = jp.GetString("type", "");
      uint32_t material = static_cast<uint32_t>(jp.GetInt("material", 0));
      PrimDef def;
      if (type == "quad") {
        V3 q, u, v;
        ReadV3(jp, "q", V3{0, 0, 0}, &q);
        ReadV3(jp, "u", V3{1, 0, 0}, &u);
        ReadV3(jp, "v", V3{0, 0, 1}, &v);
        def.refs.push_back(b.AddQuad(q, u, v, material, &def.ref_box));
      } else if (type == "box") {
        V3 a, bb;
        ReadV3(jp, "a", V3{0, 0, 0}, &a);
        ReadV3(jp, "b", V3{1, 1, 1}, &bb);
        b.AddBox(a, bb, material, &def);
      } else if (type == "sphere") {
        V3 c, disp;
        ReadV3(jp, "center", V3{0, 0, 0}, &c);
        ReadV3(jp, "displacement", V3{0, 0, 0}, &disp);
        float radius = static_cast<float>(jp.GetDouble("radius", 0.5));
        def.refs.push_back(b.AddSphere(c, disp, radius, material, &def.ref_box));
      } else {
        b.Warn("invalid primitive type '" + type + "' (entry skipped; later primitive indices shift)");
        continue;
      }
      if (!finish_medium(jp, &def)) continue;
      b.prims.push_back(std::move(def));
    }
    // scene nodes (Serialize.cpp:344-346)
    if (const json::Value* nodes = obj.Find("scene"); nodes && nodes->IsArray()) {
      for (size_t i = 0; i < nodes->Size(); i++) {
        Node n;
        int rc = ParseNode(nodes->At(i), b.prims.size(), &n, err);
        if (rc != RT2_OK) return rc;
        b.top.push_back(std::move(n));
      }
    }
  } else if (legacy) {
    // Legacy adapter (13 files in data/; HEAD's loader throws type_error.306 on these):
    // every sphere / quad / box becomes one top-level node, material_id -> material (ids equal array positions).
    b.Warn("legacy scene format: adapted (each primitive becomes a top-level scene node)");
    auto mat_of = [&](const json::Value& jp) -> uint32_t {
      int id = jp.GetInt("material_id", jp.GetInt("material", 0));
      for (size_t i = 0; i < mats->Size(); i++)
        if (mats->At(i).GetInt("id", static_cast<int>(i)) == id) return static_cast<uint32_t>(i);
      return static_cast<uint32_t>(id);
    };
    if (const json::Value* sp = jprims->Find("spheres"); sp && sp->IsArray()) {
      for (size_t i = 0; i < sp->Size(); i++) {
        const json::Value& jp = sp->At(i);
        PrimDef def;
        V3 c, disp;
        ReadV3(jp, "center", V3{0, 0, 0}, &c);
        ReadV3(jp, "displacement", V3{0, 0, 0}, &disp);
        float radius = static_cast<float>(jp.GetDouble("radius", 0.5));
        def.refs.push_back(b.AddSphere(c, disp, radius, mat_of(jp), &def.ref_box));
        if (!finish_medium(jp, &def)) continue;
        b.prims.push_back(std::move(def));
      }
    }
    if (const json::Value* qs = jprims->Find("quads"); qs && qs->IsArray()) {
      for (size_t i = 0; i < qs->Size(); i++) {
        const json::Value& jp = qs->At(i);
        PrimDef def;
        V3 q, u, v;
        ReadV3(jp, "q", V3{0, 0, 0}, &q);
        ReadV3(jp, "u", V3{1, 0, 0}, &u);
        ReadV3(jp, "v", V3{0, 0, 1}, &v);
        def.refs.push_back(b.AddQuad(q, u, v, mat_of(jp), &def.ref_box));
        if (!finish_medium(jp, &def)) continue;
        b.prims.push_back(std::move(def));
      }
    }
    if (const json::Value* bx = jprims->Find("boxes"); bx && bx->IsArray()) {
      for (size_t i = 0; i < bx->Size(); i++) {
        const json::Value& jp = bx->At(i);
        PrimDef def;
        V3 a, bb;
        ReadV3(jp, "a", V3{0, 0, 0}, &a);
        ReadV3(jp, "b", V3{1, 1, 1}, &bb);
        b.AddBox(a, bb, mat_of(jp), &def);
        if (!finish_medium(jp, &def)) continue;
        b.prims.push_back(std::move(def));
      }
    }
    for (size_t i = 0; i < b.prims.size(); i++) {
      Node n;
      n.primitive = static_cast<int>(i);
      b.top.push_back(std::move(n));
    }
  }

  // validate indices the device will dereference (the reference would read out of range)
  for (const rt2_material& m : sc->materials) {
    if ((m.type == RT2_MAT_TEXTURE || m.type == RT2_MAT_DIFFUSE_LIGHT || m.type == RT2_MAT_ISOTROPIC) &&
        m.tex_idx >= sc->textures.size()) {
      *err = "material references texture index out of range";
      return RT2_ERR_PARSE;
    }
  }
  for (const rt2_texture& t : sc->textures) {
    if (t.type == RT2_TEX_CHECKER && (t.even_tex_idx >= sc->textures.size() || t.odd_tex_idx >= sc->textures.size())) {
      *err = "checker texture references texture index out of range";
      return RT2_ERR_PARSE;
    }
  }
  auto check_mat = [&](uint32_t m) { return m < sc->materials.size(); };
  for (const rt2_sphere& s : sc->spheres)
    if (!check_mat(s.material)) {
      *err = "sphere references material index out of range";
      return RT2_ERR_PARSE;
    }
  for (const rt2_quad& q : sc->quads)
    if (!check_mat(q.material)) {
      *err = "quad references material index out of range";
      return RT2_ERR_PARSE;
    }
  for (const PrimDef& d : b.prims)
    if (d.is_medium && !check_mat(d.medium_material)) {
      *err = "constant_medium references material index out of range";
      return RT2_ERR_PARSE;
    }

  // dims (Serialize.cpp:349-357, App.cpp:115,122-125)
  sc->width = 1600;
  sc->height = 900;
  if (cam_is_object) {
    int width = cam->GetInt("width", 0);
    float aspect = cam->GetFloat("aspect_ratio", 0.0f);
    if (width != 0 && aspect != 0.0f) {
      float height = static_cast<float>(width) / aspect;
      int h = static_cast<int>(height);
      if (h != 0) {
        sc->width = width;
        sc->height = h;
      }
    }
  }
  return Compile(b, err);
}

}  // namespace

// Camera::Update (Camera.hpp:16-48), float arithmetic in the reference's order
void HostScene::UpdateCamera() {
  const float W = static_cast<float>(width), H = static_cast<float>(height);
  float theta = Radians(cam.vfov);
  float h = std::tan(theta / 2);
  V3 w = Normalize(cam.center - cam.look_at);
  V3 u = Normalize(Cross(cam.view_up, w));
  V3 v = Cross(w, u);
  float viewport_height = static_cast<float>(2.0 * static_cast<double>(h) * static_cast<double>(cam.focus_dist));
  float viewport_width = viewport_height * (W / H);
  V3 vu = viewport_width * u;
  V3 vv = viewport_height * v;
  V3 du = vu / W;
  V3 dv = vv / H;
  V3 upper_left = cam.center - (w * cam.focus_dist) - vu / 2.0f - vv / 2.0f;
  V3 p00 = upper_left + 0.5f * (du + dv);
  float defocus_radius = cam.focus_dist * std::tan(Radians(cam.defocus_angle / 2));
  V3 ddu = u * defocus_radius, ddv = v * defocus_radius;
  rt2_camera& c = camera_block;
  for (int k = 0; k < 3; k++) {
    c.center[k] = cam.center[k];
    c.pixel00[k] = p00[k];
    c.pixel_delta_u[k] = du[k];
    c.pixel_delta_v[k] = dv[k];
    c.defocus_disk_u[k] = ddu[k];
    c.defocus_disk_v[k] = ddv[k];
    c.look_at[k] = cam.look_at[k];
  }
  c.defocus_angle = cam.defocus_angle;
  c.vfov = cam.vfov;
  c.focus_dist = cam.focus_dist;
}

void HostScene::FillDesc(rt2_scene_desc* d) const {
  std::memset(d, 0, sizeof(*d));
  d->n_spheres = static_cast<uint32_t>(spheres.size());
  d->n_quads = static_cast<uint32_t>(quads.size());
  d->n_xforms = static_cast<uint32_t>(xforms.size());
  d->n_instances = static_cast<uint32_t>(instances.size());
  d->n_media = static_cast<uint32_t>(media.size());
  d->n_materials = static_cast<uint32_t>(materials.size());
  d->n_textures = static_cast<uint32_t>(textures.size());
  d->n_perlin = static_cast<uint32_t>(perlin.size());
  d->n_prim_refs = static_cast<uint32_t>(prim_refs.size());
  d->n_node_pairs = static_cast<uint32_t>(nodes.size() / 2);
  d->tlas_root = tlas_root;
  d->n_top_level = n_top_level;
  d->spheres = spheres.data();
  d->quads = quads.data();
  d->xforms = xforms.data();
  d->instances = instances.data();
  d->media = media.data();
  d->materials = materials.data();
  d->textures = textures.data();
  d->perlin = perlin.data();
  d->prim_refs = prim_refs.data();
  d->nodes = nodes.data();
  for (int k = 0; k < 3; k++) d->background[k] = background[k];
  d->min_inv_scale = min_inv_scale;
  d->width = width;
  d->height = height;
  d->camera = camera_block;
  d->n_images = static_cast<uint32_t>(images.size());
  d->n_image_texels = static_cast<uint32_t>(image_texels.size() / 4);
  d->images = images.data();
  d->image_texels = image_texels.data();
  d->has_unified_tlas = has_unified_tlas ? 1u : 0u;
  d->tlas_unified_root = tlas_unified_root;
  d->n_inst_leaves = static_cast<uint32_t>(inst_leaves.size() / 2);
  d->pad_unified = 0;
  d->inst_leaves = inst_leaves.data();
  d->has_world_tlas = has_world_tlas ? 1u : 0u;
  d->tlas_world_root = tlas_world_root;
  d->inst_bounds = inst_bounds.data();
}

int LoadSceneString(const std::string& text, const std::string& data_dir, uint64_t perlin_seed, HostScene* out,
                    std::string* err) {
  json::Value v;
  std::string perr;
  if (!json::Parser(text).Parse(v, &perr)) {
    *err = "JSON parse error: " + perr;
    return RT2_ERR_PARSE;
  }
  *out = HostScene{};
  return ParseDocument(v, data_dir, perlin_seed, false, out, err);
}

int LoadSceneFile(const std::string& path, const std::string& data_dir_in, uint64_t perlin_seed, HostScene* out,
                  std::string* err) {
  std::string text;
  if (!ReadFile(path, &text)) {
    *err = "Failed to open json file: " + path;
    return RT2_ERR_IO;
  }
  std::string data_dir = data_dir_in;
  size_t slash = path.find_last_of('/');
  std::string base = (slash == std::string::npos) ? path : path.substr(slash + 1);
  if (data_dir.empty()) data_dir = (slash == std::string::npos) ? "." : path.substr(0, slash);
  json::Value v;
  std::string perr;
  if (!json::Parser(text).Parse(v, &perr)) {
    *err = "JSON parse error in " + path + ": " + perr;
    return RT2_ERR_PARSE;
  }
  *out = HostScene{};
  bool is_final = base.rfind("final_render", 0) == 0;
  return ParseDocument(v, data_dir, perlin_seed, is_final, out, err);
}

// LoadAppSettings (Serialize.cpp:56-65).  A missing file makes the reference throw; we return RT2_ERR_IO.
int LoadAppSettings(const std::string& path, AppSettings* out, std::string* err) {
  std::string text;
  if (!ReadFile(path, &text)) {
    *err = "Failed to open json file: " + path;
    return RT2_ERR_IO;
  }
  json::Value v;
  std::string perr;
  if (!json::Parser(text).Parse(v, &perr) || !v.IsObject()) {
    *err = "settings file is not a JSON object: " + path;
    return RT2_ERR_PARSE;
  }
  out->num_samples = static_cast<size_t>(v.GetInt("num_samples", 1));
  out->render_once = v.GetBool("render_once", false);
  out->save_after_render_once = v.GetBool("save_after_render_once", false);
  out->max_depth = static_cast<size_t>(v.GetInt("max_depth", 50));
  out->render_window = v.GetBool("render_window", true);
  return RT2_OK;
}

// Synthetic BVH stress scene, SURVEY §8d config C5: n spheres, centres ~U([-1000,1000] x [0,200] x [-1000,1000]),
// radii ~U(0.2,1.0)*(1e6/n)^(1/3), ground sphere r=1e5, materials 80 % lambertian (albedo = xi*xi) / 15 % metal
// (albedo U(0.5,1), fuzz U(0,0.5)) / 5 % dielectric 1.5, background (0.7,0.8,1.0), camera (0,600,-2200) -> origin,
// fov 40, no defocus.  Every sphere is a top-level object (no transforms, no media).
int MakeSyntheticSpheres(uint32_t n, uint64_t seed, int width, int height, bool build_host_bvh, HostScene* sc, std::string* err) {
  if (n == 0 || n > 0x0FFFFFF0u) {
    *err = "sphere count out of range";
    return RT2_ERR_INVALID_ARG;
  }
  *sc = HostScene{};
  SplitMix64 rng(seed);
  sc->background[0] = 0.7f;
  sc->background[1] = 0.8f;
  sc->background[2] = 1.0f;
  sc->cam.center = V3{0, 600, -2200};
  sc->cam.look_at = V3{0, 0, 0};
  sc->cam.vfov = 40;
  sc->cam.defocus_angle = 0;
  sc->cam.focus_dist = 10;
  sc->width = width > 0 ? width : 3840;
  sc->height = height > 0 ? height : 2160;
  std::vector<BuildPrim> tlas;
  tlas.reserve(n + 1);
  sc->spheres.reserve(n + 1);
  sc->materials.reserve(n + 1);
  auto add = [&](V3 c, float r, const rt2_material& m) {
    rt2_sphere s{};
    for (int k = 0; k < 3; k++) s.center0[k] = c[k];
    s.radius = r;
    s.material = static_cast<uint32_t>(sc->materials.size());
    sc->materials.push_back(m);
    sc->spheres.push_back(s);
    uint32_t ref = (RT2_PRIM_SPHERE << 28) | static_cast<uint32_t>(sc->spheres.size() - 1);
    tlas.push_back(Flattener::MakeBuildPrim(PrimBounds(*sc, ref), ref));
  };
  rt2_material ground{};
  ground.type = RT2_MAT_LAMBERTIAN;
  ground.albedo[0] = ground.albedo[1] = ground.albedo[2] = 0.5f;
  add(V3{0, -100000.f, 0}, 100000.f, ground);
  const float rscale = std::cbrt(1.0e6f / static_cast<float>(n));
  for (uint32_t i = 0; i < n; i++) {
    V3 c{rng.Real(-1000.f, 1000.f), rng.Real(0.f, 200.f), rng.Real(-1000.f, 1000.f)};
    float r = rng.Real(0.2f, 1.0f) * rscale;
    float choose = rng.Real();
    rt2_material m{};
    if (choose < 0.8f) {
      m.type = RT2_MAT_LAMBERTIAN;
      for (int k = 0; k < 3; k++) m.albedo[k] = rng.Real() * rng.Real();
    } else if (choose < 0.95f) {
      m.type = RT2_MAT_METAL;
      for (int k = 0; k < 3; k++) m.albedo[k] = rng.Real(0.5f, 1.0f);
      m.fuzz = rng.Real(0.f, 0.5f);
    } else {
      m.type = RT2_MAT_DIELECTRIC;
      m.refraction_index = 1.5f;
    }
    add(c, r, m);
  }
  sc->n_top_level = n + 1;
  sc->span1_flags.assign(sc->n_top_level, 0);  // no media: Q2 is irrelevant
  sc->has_host_bvh = build_host_bvh;
  if (build_host_bvh) {
    sc->tlas_root = BuildBVH(tlas, sc);
  } else {
    sc->tlas_root = 0;  // the tree is built on the device (RT2_FLAG_GPU_LBVH)
  }
  sc->tree_prims.resize(1);
  sc->tree_prims[0] = std::move(tlas);
  sc->UpdateCamera();
  return RT2_OK;
}

}  // namespace rt2
