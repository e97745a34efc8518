// oracle/rt_oracle.cpp — TEST INFRASTRUCTURE: CPU restatement ("port") of the reference's path-tracing hot path.
//
// Plain C++17, no third-party code.  It re-states, in the reference's own structure (a tree of hittables walked
// recursively, closest-hit by shrinking the interval), the algorithm of tonadr1022/Raytrace2's src/cpu_raytrace; every
// function cites the reference file:line it follows.  It exists so the CUDA path can be checked on machines where the
// reference itself is not available: it is pinned against the real reference (oracle/_ref, built from the unmodified
// sources) by tests/test_oracle_vs_reference.py — bit-exact on deterministic Hit() queries, statistically on renders.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library; the product never does.
//
// The scene graph is assembled through the C calls at the bottom (driven by oracle/rt_oracle.py, which restates
// Serialize.cpp's LoadScene in Python), mirroring the reference's constructors one-to-one.
// Build: oracle/Makefile `make port` (g++ -O3 -ffp-contract=off: separate IEEE mul/add, like the reference build).
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <random>
#include <thread>
#include <vector>

namespace orc {

using real = float;                                              // Defs.hpp:16
constexpr real kInf = std::numeric_limits<real>::max();          // Defs.hpp:17 (FLT_MAX, not IEEE inf)

struct V3 {
  real x{0}, y{0}, z{0};
  real& operator[](int i) { return (&x)[i]; }
  const real& operator[](int i) const { return (&x)[i]; }
};
static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
static inline V3 operator*(V3 a, real s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(real s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
static inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 operator/(V3 a, real s) { return {a.x / s, a.y / s, a.z / s}; }
// GLM operation order: dot = (x*x' + y*y') + z*z'; normalize = v * (1/sqrt(dot)); cross as published.
static inline real dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
static inline V3 normalize(V3 v) { return v * (real(1) / std::sqrt(dot(v, v))); }
static inline real length(V3 v) { return std::sqrt(dot(v, v)); }

// column-major 4x4, m[col][row]
struct M4 {
  real m[4][4];
};
static M4 Identity() {
  M4 r{};
  for (int i = 0; i < 4; i++) r.m[i][i] = 1;
  return r;
}
// glm mat4*vec4: (m0*x + m1*y) + (m2*z + m3*w)
static inline void MulV4(const M4& a, const real v[4], real out[4]) {
  for (int r = 0; r < 4; r++) out[r] = (a.m[0][r] * v[0] + a.m[1][r] * v[1]) + (a.m[2][r] * v[2] + a.m[3][r] * v[3]);
}
static inline V3 MulPoint(const M4& a, V3 p) {
  real v[4] = {p.x, p.y, p.z, 1}, o[4];
  MulV4(a, v, o);
  return {o[0], o[1], o[2]};
}
// glm mat3*vec3 on the upper-left block: (m0*x + m1*y) + m2*z
static inline V3 MulDir(const M4& a, V3 d) {
  return {(a.m[0][0] * d.x + a.m[1][0] * d.y) + a.m[2][0] * d.z, (a.m[0][1] * d.x + a.m[1][1] * d.y) + a.m[2][1] * d.z,
          (a.m[0][2] * d.x + a.m[1][2] * d.y) + a.m[2][2] * d.z};
}
static M4 MulM(const M4& a, const M4& b) {
  M4 r;
  for (int j = 0; j < 4; j++)
    for (int i = 0; i < 4; i++) r.m[j][i] = ((a.m[0][i] * b.m[j][0] + a.m[1][i] * b.m[j][1]) + a.m[2][i] * b.m[j][2]) + a.m[3][i] * b.m[j][3];
  return r;
}
// glm::inverse(mat4) by cofactors
static M4 Inverse(const M4& mm) {
  const auto& m = mm.m;
  real c00 = m[2][2] * m[3][3] - m[3][2] * m[2][3], c02 = m[1][2] * m[3][3] - m[3][2] * m[1][3], c03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
  real c04 = m[2][1] * m[3][3] - m[3][1] * m[2][3], c06 = m[1][1] * m[3][3] - m[3][1] * m[1][3], c07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
  real c08 = m[2][1] * m[3][2] - m[3][1] * m[2][2], c10 = m[1][1] * m[3][2] - m[3][1] * m[1][2], c11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
  real c12 = m[2][0] * m[3][3] - m[3][0] * m[2][3], c14 = m[1][0] * m[3][3] - m[3][0] * m[1][3], c15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
  real c16 = m[2][0] * m[3][2] - m[3][0] * m[2][2], c18 = m[1][0] * m[3][2] - m[3][0] * m[1][2], c19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
  real c20 = m[2][0] * m[3][1] - m[3][0] * m[2][1], c22 = m[1][0] * m[3][1] - m[3][0] * m[1][1], c23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
  real f0[4] = {c00, c00, c02, c03}, f1[4] = {c04, c04, c06, c07}, f2[4] = {c08, c08, c10, c11};
  real f3[4] = {c12, c12, c14, c15}, f4[4] = {c16, c16, c18, c19}, f5[4] = {c20, c20, c22, c23};
  real v0[4] = {m[1][0], m[0][0], m[0][0], m[0][0]}, v1[4] = {m[1][1], m[0][1], m[0][1], m[0][1]};
  real v2[4] = {m[1][2], m[0][2], m[0][2], m[0][2]}, v3[4] = {m[1][3], m[0][3], m[0][3], m[0][3]};
  const real sa[4] = {1, -1, 1, -1}, sb[4] = {-1, 1, -1, 1};
  M4 inv;
  for (int i = 0; i < 4; i++) {
    inv.m[0][i] = ((v1[i] * f0[i] - v2[i] * f1[i]) + v3[i] * f2[i]) * sa[i];
    inv.m[1][i] = ((v0[i] * f0[i] - v2[i] * f3[i]) + v3[i] * f4[i]) * sb[i];
    inv.m[2][i] = ((v0[i] * f1[i] - v1[i] * f3[i]) + v3[i] * f5[i]) * sa[i];
    inv.m[3][i] = ((v0[i] * f2[i] - v1[i] * f4[i]) + v2[i] * f5[i]) * sb[i];
  }
  real d0 = m[0][0] * inv.m[0][0], d1 = m[0][1] * inv.m[1][0], d2 = m[0][2] * inv.m[2][0], d3 = m[0][3] * inv.m[3][0];
  real det = (d0 + d1) + (d2 + d3);
  real ood = real(1) / det;
  M4 r;
  for (int j = 0; j < 4; j++)
    for (int i = 0; i < 4; i++) r.m[j][i] = inv.m[j][i] * ood;
  return r;
}
// Serialize.cpp:106-132: translate(I, t) * toMat4(angleAxis(radians(angle), axis)) * scale(I, s)
static M4 ComposeTRS(const real t[3], int has_rot, const real aa[4], const real s[3]) {
  M4 T = Identity();
  for (int i = 0; i < 4; i++) T.m[3][i] = ((T.m[0][i] * t[0] + T.m[1][i] * t[1]) + T.m[2][i] * t[2]) + T.m[3][i];
  real qw = 1, qx = 0, qy = 0, qz = 0;  // absent rotation: identity (the reference leaves the quat uninitialised, :114)
  if (has_rot) {
    real ang = aa[0] * static_cast<real>(0.01745329251994329576923690768489);
    real sn = std::sin(ang * real(0.5));
    qw = std::cos(ang * real(0.5));
    qx = aa[1] * sn, qy = aa[2] * sn, qz = aa[3] * sn;
  }
  M4 R = Identity();
  real qxx = qx * qx, qyy = qy * qy, qzz = qz * qz, qxz = qx * qz, qxy = qx * qy, qyz = qy * qz, qwx = qw * qx, qwy = qw * qy, qwz = qw * qz;
  R.m[0][0] = 1 - 2 * (qyy + qzz), R.m[0][1] = 2 * (qxy + qwz), R.m[0][2] = 2 * (qxz - qwy);
  R.m[1][0] = 2 * (qxy - qwz), R.m[1][1] = 1 - 2 * (qxx + qzz), R.m[1][2] = 2 * (qyz + qwx);
  R.m[2][0] = 2 * (qxz + qwy), R.m[2][1] = 2 * (qyz - qwx), R.m[2][2] = 1 - 2 * (qxx + qyy);
  M4 S = Identity();
  for (int i = 0; i < 4; i++) S.m[0][i] *= s[0], S.m[1][i] *= s[1], S.m[2][i] *= s[2];
  return MulM(MulM(T, R), S);
}

// ---- RNG: Math.hpp:9-43 ------------------------------------------------------------------------------------------
static thread_local std::minstd_rand tl_gen{std::random_device{}()};
static inline real RandReal() {
  static thread_local std::uniform_real_distribution<real> dist(0.0, 1.0);
  return dist(tl_gen);
}
static inline real RandReal(real lo, real hi) { return lo + RandReal() * (hi - lo); }
static inline int RandInt(int lo, int hi) { return static_cast<int>(RandReal(static_cast<real>(lo), static_cast<real>(hi + 1))); }
static inline V3 RandInUnitSphere() {
  while (true) {
    V3 p{RandReal(-1, 1), RandReal(-1, 1), RandReal(-1, 1)};
    real l2 = dot(p, p);
    if (1e-160 < l2 && l2 <= 1.0) return p;
  }
}
static inline V3 RandInUnitDisk() {
  while (true) {
    V3 p{RandReal(-1, 1), RandReal(-1, 1), 0};
    if (dot(p, p) < 1.0) return p;
  }
}
static inline V3 RandUnitVec3() { return normalize(RandInUnitSphere()); }
static inline bool NearZero(V3 v) { return std::fabs(v.x) < 1e-8 && std::fabs(v.y) < 1e-8 && std::fabs(v.z) < 1e-8; }  // Math.hpp:61-64
static inline V3 Reflect(V3 v, V3 n) { return v - real(2) * dot(v, n) * n; }                                               // Math.hpp:66
static inline V3 Refract(V3 uv, V3 n, real eta) {                                                                           // Math.hpp:68-73
  real cos_theta = static_cast<real>(std::fmin(static_cast<double>(dot(-uv, n)), 1.0));
  V3 perp = eta * (uv + cos_theta * n);
  V3 par = -std::sqrt(std::fabs(1.0f - dot(perp, perp))) * n;
  return perp + par;
}

// ---- Ray / Interval / AABB / HitRecord (Ray.hpp, Interval.hpp, AABB.hpp, HitRecord.hpp) ---------------------------------
struct Ray {
  V3 o, d;
  real time{0};
  V3 At(real t) const { return o + d * t; }
};
struct Interval {
  real min{kInf}, max{-kInf};
  Interval() = default;
  Interval(real a, real b) : min(a), max(b) {}
  Interval(const Interval& a, const Interval& b) : min(std::fmin(a.min, b.min)), max(std::fmax(a.max, b.max)) {}
  real Size() const { return max - min; }
  bool Contains(real x) const { return min <= x && x <= max; }
  bool Surrounds(real x) const { return min < x && x < max; }
  Interval Expand(real delta) const {
    real pad = delta / 2.0f;
    return {min - pad, max + pad};
  }
};
struct AABB {
  Interval x, y, z;
  AABB() = default;
  AABB(V3 a, V3 b) : x(std::fmin(a.x, b.x), std::fmax(a.x, b.x)), y(std::fmin(a.y, b.y), std::fmax(a.y, b.y)), z(std::fmin(a.z, b.z), std::fmax(a.z, b.z)) { Pad(); }
  AABB(const AABB& a, const AABB& b) : x(a.x, b.x), y(a.y, b.y), z(a.z, b.z) { Pad(); }
  const Interval& Axis(int n) const { return n == 0 ? x : (n == 1 ? y : z); }
  // AABB.hpp:34-47
  bool Hit(const Ray& r, Interval t) const {
    for (int a = 0; a < 3; a++) {
      const Interval& ax = Axis(a);
      const real inv = 1.f / r.d[a];
      real t0 = (ax.min - r.o[a]) * inv, t1 = (ax.max - r.o[a]) * inv;
      if (t1 < t0) std::swap(t0, t1);
      t.min = (t0 < t.min) ? t.min : t0;  // glm::max(t0, t.min)
      t.max = (t.max < t1) ? t.max : t1;  // glm::min(t1, t.max)
      if (t.max <= t.min) return false;
    }
    return true;
  }
  int LongestAxis() const {
    if (x.Size() > y.Size()) return x.Size() > z.Size() ? 0 : 2;
    return y.Size() > z.Size() ? 1 : 2;
  }
  void Pad() {
    constexpr real kDelta = 0.0001f;
    if (x.Size() < kDelta) x = x.Expand(kDelta);
    if (y.Size() < kDelta) y = y.Expand(kDelta);
    if (z.Size() < kDelta) z = z.Expand(kDelta);
  }
};
struct HitRecord {
  V3 point, normal;
  real u{0}, v{0};  // HitRecord::uv (HitRecord.hpp:12): set by Sphere::Hit / Quad::IsInterior, read only by image textures
  real t{0};
  int material{-1};
  bool front_face{false};
  int leaf_id{-1};  // restatement extra: which leaf primitive produced the record
  void SetFaceNormal(const Ray& r, V3 outward) {  // HitRecord.hpp:17-20
    front_face = dot(r.d, outward) < 0;
    normal = (static_cast<real>(static_cast<int>(front_face) << 1) - 1.0f) * outward;
  }
};

struct Scene;
struct Hittable {
  virtual ~Hittable() = default;
  virtual bool Hit(const Scene& s, const Ray& r, Interval t, HitRecord& rec) const = 0;
  virtual AABB Box() const = 0;
};
using HP = std::shared_ptr<Hittable>;

// Sphere.hpp:14-34, Sphere.cpp:7-37
struct Sphere : Hittable {
  Ray center;  // origin = start, d = displacement
  real radius;
  int material, leaf_id;
  AABB box;
  Sphere(V3 c, V3 disp, real r, int mat, int id) : radius(r), material(mat), leaf_id(id) {
    center.o = c, center.d = disp;
    V3 rv{r, r, r};
    box = AABB{AABB{center.At(0) - rv, center.At(0) + rv}, AABB{center.At(1) - rv, center.At(1) + rv}};
  }
  bool Hit(const Scene&, const Ray& r, Interval t, HitRecord& rec) const override {
    V3 cc = center.At(r.time);
    V3 oc = cc - r.o;
    real a = dot(r.d, r.d), h = dot(r.d, oc), c = dot(oc, oc) - radius * radius;
    real disc = h * h - a * c;
    if (disc < 0) return false;
    real sq = std::sqrt(disc);
    real root = (h - sq) / a;
    if (!t.Surrounds(root)) {
      root = (h + sq) / a;
      if (!t.Surrounds(root)) return false;
    }
    rec.t = root;
    rec.point = r.At(root);
    rec.material = material;
    rec.leaf_id = leaf_id;
    V3 outward = (rec.point - cc) / radius;
    rec.SetFaceNormal(r, outward);
    {  // rec.uv = GetUV(outward_normal) (Sphere.cpp:34,39-43)
      const real pi = 3.14159265358979323846f;
      real theta = std::acos(-outward.y), phi = std::atan2(-outward.z, outward.x) + pi;
      rec.u = phi / (2.f * pi);
      rec.v = theta / pi;
    }
    return true;
  }
  AABB Box() const override { return box; }
};

// Quad.hpp:13-32, Quad.cpp:8-43
struct Quad : Hittable {
  V3 q, u, v, w, normal;
  real d;
  int material, leaf_id;
  AABB box;
  Quad(V3 q_, V3 u_, V3 v_, int mat, int id) : q(q_), u(u_), v(v_), material(mat), leaf_id(id) {
    V3 n = cross(u, v);
    normal = normalize(n);
    d = dot(normal, q);
    w = n / dot(n, n);
    box = AABB{AABB{q, q + u + v}, AABB{q + u, q + v}};
  }
  bool Hit(const Scene&, const Ray& r, Interval t, HitRecord& rec) const override {
    real ndd = dot(normal, r.d);
    if (std::fabs(ndd) < 1e-8) return false;
    real tt = (d - dot(normal, r.o)) / ndd;
    if (!t.Contains(tt)) return false;
    V3 p = r.At(tt);
    V3 ph = p - q;
    real alpha = dot(w, cross(ph, v)), beta = dot(w, cross(u, ph));
    Interval unit{0, 1};
    if (!unit.Contains(alpha) || !unit.Contains(beta)) return false;
    rec.t = tt;
    rec.point = p;
    rec.u = alpha, rec.v = beta;  // Quad.cpp:15
    rec.material = material;
    rec.leaf_id = leaf_id;
    rec.SetFaceNormal(r, normal);
    return true;
  }
  AABB Box() const override { return box; }
};

// HittableList.hpp:8-23, HittableList.cpp:8-22
struct List : Hittable {
  std::vector<HP> objs;
  AABB box;
  void Add(const HP& o) {
    objs.push_back(o);
    box = AABB{box, o->Box()};
  }
  bool Hit(const Scene& s, const Ray& r, Interval t, HitRecord& rec) const override {
    HitRecord tmp;
    bool any = false;
    for (const HP& o : objs) {
      if (o->Hit(s, r, t, tmp)) {
        any = true;
        t.max = tmp.t;
        rec = tmp;
      }
    }
    return any;
  }
  AABB Box() const override { return box; }
};

// Transform.hpp:20-36, Transform.cpp:13-20,36-64,75-88
struct Transformed : Hittable {
  M4 model, inv;
  HP obj;
  AABB box;
  Transformed(HP o, const M4& m) : model(m), obj(std::move(o)) {
    inv = Inverse(model);
    AABB b = obj->Box();
    V3 mn{b.x.min, b.y.min, b.z.min}, mx{b.x.max, b.y.max, b.z.max};
    V3 corners[8] = {{mn.x, mn.y, mn.z}, {mx.x, mn.y, mn.z}, {mn.x, mx.y, mn.z}, {mx.x, mx.y, mn.z},
                     {mn.x, mn.y, mx.z}, {mx.x, mn.y, mx.z}, {mn.x, mx.y, mx.z}, {mx.x, mx.y, mx.z}};
    V3 nmin{kInf, kInf, kInf}, nmax{-kInf, -kInf, -kInf};
    for (const V3& c : corners) {
      V3 tc = MulPoint(model, c);
      for (int k = 0; k < 3; k++) nmin[k] = std::fmin(nmin[k], tc[k]), nmax[k] = std::fmax(nmax[k], tc[k]);
    }
    box = AABB{nmin, nmax};
  }
  bool Hit(const Scene& s, const Ray& r, Interval t, HitRecord& rec) const override {
    // model ray: direction NORMALISED, ray_t passed through unchanged, rec.t never rescaled (quirk Q1)
    Ray m{MulPoint(inv, r.o), normalize(MulDir(inv, r.d)), r.time};
    if (!obj->Hit(s, m, t, rec)) return false;
    rec.point = MulPoint(model, rec.point);
    // normal_mat = mat3(transpose(inverse(model))): (N*n)_r = (inv[r][0]*nx + inv[r][1]*ny) + inv[r][2]*nz in [col][row] storage
    V3 n = rec.normal;
    V3 nn{(inv.m[0][0] * n.x + inv.m[0][1] * n.y) + inv.m[0][2] * n.z, (inv.m[1][0] * n.x + inv.m[1][1] * n.y) + inv.m[1][2] * n.z,
          (inv.m[2][0] * n.x + inv.m[2][1] * n.y) + inv.m[2][2] * n.z};
    rec.normal = normalize(nn);
    return true;
  }
  AABB Box() const override { return box; }
};

// ConstantMedium.hpp:5-16, ConstantMedium.cpp:10-58
struct Medium : Hittable {
  HP boundary;
  real neg_inv_density;
  int material, leaf_id;
  Medium(HP b, real density, int mat, int id) : boundary(std::move(b)), neg_inv_density(static_cast<real>(-1.0 / density)), material(mat), leaf_id(id) {}
  bool Hit(const Scene& s, const Ray& r, Interval t, HitRecord& rec) const override {
    HitRecord r1, r2;
    if (!boundary->Hit(s, r, Interval{-kInf, kInf}, r1)) return false;
    if (!boundary->Hit(s, r, Interval(static_cast<real>(r1.t + 0.0001), kInf), r2)) return false;
    r1.t = std::fmax(r1.t, t.min);
    r2.t = std::fmin(r2.t, t.max);
    if (r1.t >= r2.t) return false;
    r1.t = std::fmax(r1.t, 0.0f);
    real len = length(r.d);
    real inside = (r2.t - r1.t) * len;
    real hit_dist = neg_inv_density * std::log(RandReal());
    if (hit_dist > inside) return false;
    rec.t = r1.t + hit_dist / len;
    rec.point = r.At(rec.t);
    rec.normal = V3{1, 0, 0};
    rec.front_face = true;
    rec.material = material;
    rec.leaf_id = leaf_id;
    return true;
  }
  AABB Box() const override { return boundary->Box(); }
};

// BVH.hpp:11-28, BVH.cpp:10-55
struct BVHNode : Hittable {
  HP left, right;
  AABB box;
  BVHNode(std::vector<HP>& objs, size_t start, size_t end) {
    for (size_t i = start; i < end; i++) box = AABB{box, objs[i]->Box()};
    size_t span = end - start;
    if (span == 1) {
      left = right = objs[start];  // the same object on both sides: Hit() runs twice (quirk Q2)
    } else if (span == 2) {
      left = objs[start];
      right = objs[start + 1];
    } else {
      int axis = box.LongestAxis();
      std::sort(objs.begin() + static_cast<long>(start), objs.begin() + static_cast<long>(end),
                [axis](const HP& a, const HP& b) { return a->Box().Axis(axis).min < b->Box().Axis(axis).min; });
      size_t mid = start + span / 2;
      left = std::make_shared<BVHNode>(objs, start, mid);
      right = std::make_shared<BVHNode>(objs, mid, end);
    }
  }
  bool Hit(const Scene& s, const Ray& r, Interval t, HitRecord& rec) const override {
    if (!box.Hit(r, t)) return false;
    bool hl = left->Hit(s, r, t, rec);
    bool hr = right->Hit(s, r, Interval{t.min, hl ? rec.t : t.max}, rec);
    return hl || hr;
  }
  AABB Box() const override { return box; }
};

// ---- textures / Perlin (Texture.hpp, Texture.cpp, PerlinNoiseGen.cpp) ------------------------------------------------
struct Perlin {
  int count{256};
  std::vector<int> px, py, pz;
  std::vector<V3> vec;
  void Init() {  // PerlinNoiseGen.cpp:41-50,90-103
    vec.resize(count);
    for (int i = 0; i < count; i++) vec[i] = normalize(V3{RandReal(-1, 1), RandReal(-1, 1), RandReal(-1, 1)});
    for (auto* p : {&px, &py, &pz}) {
      p->resize(count);
      for (int i = 0; i < count; i++) (*p)[i] = i;
      for (int i = count - 1; i > 0; i--) std::swap((*p)[i], (*p)[RandInt(0, i)]);
    }
  }
  real Noise(V3 p) const {  // PerlinNoiseGen.cpp:10-26,66-88
    real u = p.x - std::floor(p.x), v = p.y - std::floor(p.y), w = p.z - std::floor(p.z);
    int i = static_cast<int>(std::floor(p.x)), j = static_cast<int>(std::floor(p.y)), k = static_cast<int>(std::floor(p.z));
    real uu = u * u * (3 - 2 * u), vv = v * v * (3 - 2 * v), ww = w * w * (3 - 2 * w);
    real acc = 0;
    for (int di = 0; di < 2; di++)
      for (int dj = 0; dj < 2; dj++)
        for (int dk = 0; dk < 2; dk++) {
          V3 c = vec[px[(i + di) & 255] ^ py[(j + dj) & 255] ^ pz[(k + dk) & 255]];
          V3 wv{u - di, v - dj, w - dk};
          acc += (di * uu + (1 - di) * (1 - uu)) * (dj * vv + (1 - dj) * (1 - vv)) * (dk * ww + (1 - dk) * (1 - ww)) * dot(c, wv);
        }
    return acc;
  }
  real Turb(V3 p) const {  // PerlinNoiseGen.cpp:52-64, depth 7
    real acc = 0, weight = 1;
    for (int i = 0; i < 7; i++) {
      acc += weight * Noise(p);
      weight *= 0.5f;
      p = p * real(2);
    }
    return std::fabs(acc);
  }
};
struct Texture {
  int type{0};  // 0 solid, 1 checker, 2 noise, 3 image (schema extension: nearest texel of a linear RGB image, v flipped)
  int img_w{0}, img_h{0};
  std::vector<float> texels;  // 3 floats per texel, row 0 = top
  V3 albedo{1, 1, 1};
  real scale{1};  // checker: inv_scale
  int even{0}, odd{0}, noise_type{1};
  Perlin perlin;
};
struct Material {
  int type{0};  // 0 lambertian 1 metal 2 dielectric 3 texture 4 diffuse_light 5 isotropic (matches rt2.h)
  V3 albedo{0, 0, 0};
  real fuzz{0}, ior{1};
  int tex{0};
};

// Camera.hpp:16-67
struct Camera {
  V3 center{0, 0, 0}, lookat{0, 0, -1}, vup{0, 1, 0};
  real vfov{90}, defocus_angle{0}, focus{10};
  int w{1}, h{1}, spp{1};
  V3 p00, du, dv, ddu, ddv;
  int sq{1};
  real recip{1};
  void Update() {
    real theta = vfov * static_cast<real>(0.01745329251994329576923690768489);
    real hh = std::tan(theta / 2);
    V3 ww = normalize(center - lookat), uu = normalize(cross(vup, ww)), vv = cross(ww, uu);
    real vh = static_cast<real>(2.0 * hh * focus);
    real vw = vh * (static_cast<real>(w) / h);
    V3 vu = vw * uu, vvv = vh * vv;
    du = vu / static_cast<real>(w);
    dv = vvv / static_cast<real>(h);
    V3 ul = center - (ww * focus) - vu / real(2) - vvv / real(2);
    p00 = ul + real(0.5) * (du + dv);
    real dr = focus * std::tan((defocus_angle / 2) * static_cast<real>(0.01745329251994329576923690768489));
    ddu = uu * dr, ddv = vv * dr;
    sq = static_cast<int>(std::sqrt(spp));
    recip = static_cast<real>(1.0 / sq);
  }
  Ray GetRay(int x, int y, int s_i, int s_j) const {
    real ox = static_cast<real>((s_i + RandReal()) * recip - 0.5);
    real oy = static_cast<real>((s_j + RandReal()) * recip - 0.5);
    V3 pc = p00 + ((static_cast<real>(x) + ox) * du) + ((static_cast<real>(y) + oy) * dv);
    V3 c = center;
    if (!(defocus_angle <= 0)) {
      V3 p = RandInUnitDisk();
      c = center + (p.x * ddu) + (p.y * ddv);
    }
    real time = RandReal();
    return Ray{c, normalize(pc - c), time};
  }
};

struct Scene {
  std::vector<Texture> textures;
  std::vector<Material> materials;
  std::vector<HP> prims;   // JSON "primitives" (after the constant_medium wrap)
  std::vector<HP> handles; // every object made through the C API, by handle
  std::vector<HP> top;     // JSON "scene" entries
  HP root;                 // HittableList{BVHNode}, App.cpp:126
  V3 background{1, 1, 1};
  Camera cam;
  int next_leaf{0};
};

static V3 TexValue(const Scene& s, int idx, V3 p, real u = 0, real v = 0) {  // Texture.hpp:14-17, Texture.cpp:7-22
  const Texture& t = s.textures[idx];
  if (t.type == 3) {  // image_texture::value of the book the reference follows (no counterpart in the reference itself)
    if (t.img_h <= 0) return V3{0, 1, 1};
    u = std::fmin(std::fmax(u, real(0)), real(1));
    v = 1.0f - std::fmin(std::fmax(v, real(0)), real(1));
    int i = std::min(static_cast<int>(u * t.img_w), t.img_w - 1), j = std::min(static_cast<int>(v * t.img_h), t.img_h - 1);
    const float* c = &t.texels[(static_cast<size_t>(j) * t.img_w + i) * 3];
    return V3{c[0], c[1], c[2]};
  }
  if (t.type == 1) {
    int ix = static_cast<int>(std::floor(t.scale * p.x)), iy = static_cast<int>(std::floor(t.scale * p.y)), iz = static_cast<int>(std::floor(t.scale * p.z));
    return TexValue(s, (ix + iy + iz) % 2 == 0 ? t.even : t.odd, p, u, v);
  }
  if (t.type == 2) {
    if (t.noise_type == 0) return t.albedo * real(0.5) * (1.0f + t.perlin.Noise(t.scale * p));
    return t.albedo * real(0.5) * (1 + std::sin(t.scale * p.z + 10 * t.perlin.Turb(p)));
  }
  return t.albedo;
}

// Material.cpp:10-83; returns false for the emitter (base-class Scatter, Material.hpp:14-20)
static bool Scatter(const Scene& s, const Material& m, const Ray& in, const HitRecord& rec, V3& att, Ray& out) {
  switch (m.type) {
    case 0:
    case 3: {
      V3 d = rec.normal + RandUnitVec3();
      if (NearZero(d)) d = rec.normal;
      out = Ray{rec.point, d, in.time};
      att = (m.type == 0) ? m.albedo : TexValue(s, m.tex, rec.point, rec.u, rec.v);
      return true;
    }
    case 1: {
      V3 refl = normalize(Reflect(in.d, rec.normal)) + (m.fuzz * RandUnitVec3());
      out = Ray{rec.point, refl, in.time};
      att = m.albedo;
      return true;  // never absorbs (quirk Q3)
    }
    case 2: {
      att = V3{1, 1, 1};
      real ri = rec.front_face ? static_cast<real>(1.0 / m.ior) : m.ior;
      V3 ud = normalize(in.d);
      real ct = dot(-ud, rec.normal);
      ct = (real(1) < ct) ? real(1) : ct;  // glm::min(x, 1)
      real st = std::sqrt(1.f - ct * ct);
      bool cannot = ri * st > 1.0;
      real r0 = (1 - ri) / (1 + ri);
      r0 = r0 * r0;
      double schlick = r0 + (1 - r0) * std::pow(static_cast<double>(1 - ct), 5.0);  // glm::pow(float,int) -> std::pow -> double
      V3 d = (cannot || schlick > RandReal()) ? Reflect(ud, rec.normal) : Refract(ud, rec.normal, ri);
      out = Ray{rec.point, d, in.time};
      return true;
    }
    case 5: {
      out = Ray{rec.point, RandUnitVec3(), in.time};
      att = TexValue(s, m.tex, rec.point, rec.u, rec.v);
      return true;
    }
    default: return false;
  }
}

static thread_local uint64_t tl_rays = 0;
// RayTracer.cpp:20-45
static V3 RayColor(const Ray& r, int depth, const Scene& s) {
  if (depth <= 0) return {0, 0, 0};
  HitRecord rec;
  tl_rays++;
  if (!s.root->Hit(s, r, Interval{0.001f, kInf}, rec)) return s.background;
  const Material& m = s.materials[rec.material];
  V3 emit = (m.type == 4) ? TexValue(s, m.tex, rec.point, rec.u, rec.v) : V3{0, 0, 0};
  V3 att;
  Ray out;
  if (Scatter(s, m, r, rec, att, out)) return att * RayColor(out, depth - 1, s) + emit;
  return emit;
}

}  // namespace orc

using namespace orc;

extern "C" {

void* orc_scene_new() { return new Scene; }
void orc_scene_free(void* h) { delete static_cast<Scene*>(h); }

int orc_add_texture(void* h, int type, const float* albedo, float scale, int even, int odd, int noise_type, int point_count) {
  auto* s = static_cast<Scene*>(h);
  Texture t;
  t.type = type;
  t.albedo = V3{albedo[0], albedo[1], albedo[2]};
  t.scale = (type == 1) ? 1.f / scale : scale;  // Checker ctor stores inv_scale (Texture.hpp:20-21)
  t.even = even, t.odd = odd, t.noise_type = noise_type;
  if (type == 2) {
    t.perlin.count = point_count;
    t.perlin.Init();
  }
  s->textures.push_back(std::move(t));
  return static_cast<int>(s->textures.size() - 1);
}
int orc_add_material(void* h, int type, const float* albedo, float fuzz, float ior, int tex) {
  auto* s = static_cast<Scene*>(h);
  Material m;
  m.type = type;
  m.albedo = V3{albedo[0], albedo[1], albedo[2]};
  m.fuzz = fuzz, m.ior = ior, m.tex = tex;
  s->materials.push_back(m);
  return static_cast<int>(s->materials.size() - 1);
}
static int Keep(Scene* s, HP p) {
  s->handles.push_back(std::move(p));
  return static_cast<int>(s->handles.size() - 1);
}
int orc_make_sphere(void* h, const float* c, const float* disp, float radius, int material) {
  auto* s = static_cast<Scene*>(h);
  return Keep(s, std::make_shared<Sphere>(V3{c[0], c[1], c[2]}, V3{disp[0], disp[1], disp[2]}, radius, material, s->next_leaf++));
}
int orc_make_quad(void* h, const float* q, const float* u, const float* v, int material) {
  auto* s = static_cast<Scene*>(h);
  return Keep(s, std::make_shared<Quad>(V3{q[0], q[1], q[2]}, V3{u[0], u[1], u[2]}, V3{v[0], v[1], v[2]}, material, s->next_leaf++));
}
// MakeBox, Quad.hpp:34-50
int orc_make_box(void* h, const float* a, const float* b, int material) {
  auto* s = static_cast<Scene*>(h);
  V3 mn{std::fmin(a[0], b[0]), std::fmin(a[1], b[1]), std::fmin(a[2], b[2])}, mx{std::fmax(a[0], b[0]), std::fmax(a[1], b[1]), std::fmax(a[2], b[2])};
  V3 dx{mx.x - mn.x, 0, 0}, dy{0, mx.y - mn.y, 0}, dz{0, 0, mx.z - mn.z};
  auto list = std::make_shared<List>();
  auto add = [&](V3 q, V3 u, V3 v) { list->Add(std::make_shared<Quad>(q, u, v, material, s->next_leaf++)); };
  add(V3{mn.x, mn.y, mx.z}, dx, dy);
  add(V3{mx.x, mn.y, mx.z}, -dz, dy);
  add(V3{mx.x, mn.y, mn.z}, -dx, dy);
  add(V3{mn.x, mn.y, mn.z}, dz, dy);
  add(V3{mn.x, mx.y, mx.z}, dx, -dz);
  add(V3{mn.x, mn.y, mn.z}, dx, dz);
  return Keep(s, list);
}
int orc_make_medium(void* h, int boundary, float density, int material) {
  auto* s = static_cast<Scene*>(h);
  return Keep(s, std::make_shared<Medium>(s->handles[boundary], density, material, s->next_leaf++));
}
int orc_make_list(void* h) { return Keep(static_cast<Scene*>(h), std::make_shared<List>()); }
void orc_list_add(void* h, int list, int child) {
  auto* s = static_cast<Scene*>(h);
  static_cast<List*>(s->handles[list].get())->Add(s->handles[child]);
}
int orc_make_transformed(void* h, int child, const float* translation, int has_rotation, const float* angle_axis, const float* scale) {
  auto* s = static_cast<Scene*>(h);
  return Keep(s, std::make_shared<Transformed>(s->handles[child], ComposeTRS(translation, has_rotation, angle_axis, scale)));
}
void orc_add_top_level(void* h, int obj) {
  auto* s = static_cast<Scene*>(h);
  s->top.push_back(s->handles[obj]);
}
void orc_set_background(void* h, const float* bg) { static_cast<Scene*>(h)->background = V3{bg[0], bg[1], bg[2]}; }
void orc_set_camera(void* h, const float* center, const float* lookat, float vfov, float defocus_angle, float focus, int w, int hh, int spp) {
  auto* s = static_cast<Scene*>(h);
  Camera& c = s->cam;
  c.center = V3{center[0], center[1], center[2]};
  c.lookat = V3{lookat[0], lookat[1], lookat[2]};
  c.vfov = vfov, c.defocus_angle = defocus_angle, c.focus = focus;
  c.w = w, c.h = hh, c.spp = spp;
  c.Update();
}
// App.cpp:126: scene.hittable_list = HittableList{make_shared<BVHNode>(scene.hittable_list)}
void orc_finalize(void* h) {
  auto* s = static_cast<Scene*>(h);
  std::vector<HP> objs = s->top;
  auto list = std::make_shared<List>();
  list->Add(std::make_shared<BVHNode>(objs, 0, objs.size()));
  s->root = list;
}
int orc_span1(void* h, uint8_t* flags) {
  auto* s = static_cast<Scene*>(h);
  std::memset(flags, 0, s->top.size());
  std::vector<const BVHNode*> stack{static_cast<const BVHNode*>(static_cast<List*>(s->root.get())->objs[0].get())};
  while (!stack.empty()) {
    const BVHNode* n = stack.back();
    stack.pop_back();
    if (n->left == n->right) {
      for (size_t i = 0; i < s->top.size(); i++)
        if (s->top[i] == n->left) flags[i] = 1;
      continue;
    }
    for (const HP& c : {n->left, n->right})
      if (auto* b = dynamic_cast<const BVHNode*>(c.get())) stack.push_back(b);
  }
  return static_cast<int>(s->top.size());
}
void orc_camera(void* h, float* out) {
  const Camera& c = static_cast<Scene*>(h)->cam;
  const V3* v[6] = {&c.center, &c.p00, &c.du, &c.dv, &c.ddu, &c.ddv};
  for (int i = 0; i < 6; i++)
    for (int k = 0; k < 3; k++) out[i * 3 + k] = (*v[i])[k];
}
int orc_perlin_get(void* h, int tex, int32_t* px, int32_t* py, int32_t* pz, float* vec) {
  auto* s = static_cast<Scene*>(h);
  const Perlin& p = s->textures[tex].perlin;
  for (int i = 0; i < p.count; i++) {
    px[i] = p.px[i], py[i] = p.py[i], pz[i] = p.pz[i];
    for (int k = 0; k < 3; k++) vec[i * 3 + k] = p.vec[i][k];
  }
  return p.count;
}
int orc_perlin_set(void* h, int tex, const int32_t* px, const int32_t* py, const int32_t* pz, const float* vec) {
  auto* s = static_cast<Scene*>(h);
  Perlin& p = s->textures[tex].perlin;
  for (int i = 0; i < p.count; i++) {
    p.px[i] = px[i], p.py[i] = py[i], p.pz[i] = pz[i];
    p.vec[i] = V3{vec[i * 3], vec[i * 3 + 1], vec[i * 3 + 2]};
  }
  return p.count;
}
// Image texture (schema extension): rgb = 3 floats per texel, already linear, row 0 = top.
int orc_add_image_texture(void* h, int w, int ht, const float* rgb) {
  auto* s = static_cast<Scene*>(h);
  Texture t;
  t.type = 3;
  t.img_w = w, t.img_h = ht;
  t.texels.assign(rgb, rgb + static_cast<size_t>(w) * ht * 3);
  s->textures.push_back(std::move(t));
  return static_cast<int>(s->textures.size() - 1);
}
void orc_texture_value_uv(void* h, int tex, const float* pts, const float* uv, size_t n, float* rgb) {
  auto* s = static_cast<Scene*>(h);
  for (size_t i = 0; i < n; i++) {
    V3 c = TexValue(*s, tex, V3{pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]}, uv[i * 2], uv[i * 2 + 1]);
    for (int k = 0; k < 3; k++) rgb[i * 3 + k] = c[k];
  }
}
// orc_intersect + HitRecord::uv
void orc_intersect_uv(void* h, const float* rays, size_t n, float tmin, float tmax, uint8_t* hit, float* uv) {
  auto* s = static_cast<Scene*>(h);
  for (size_t i = 0; i < n; i++) {
    const float* r = rays + i * 7;
    Ray ray{V3{r[0], r[1], r[2]}, V3{r[3], r[4], r[5]}, r[6]};
    HitRecord rec;
    bool got = s->root->Hit(*s, ray, Interval{tmin, tmax}, rec);
    hit[i] = got;
    uv[i * 2] = got ? rec.u : 0;
    uv[i * 2 + 1] = got ? rec.v : 0;
  }
}
void orc_texture_value(void* h, int tex, const float* pts, size_t n, float* rgb) {
  auto* s = static_cast<Scene*>(h);
  for (size_t i = 0; i < n; i++) {
    V3 c = TexValue(*s, tex, V3{pts[i * 3], pts[i * 3 + 1], pts[i * 3 + 2]});
    rgb[i * 3] = c.x, rgb[i * 3 + 1] = c.y, rgb[i * 3 + 2] = c.z;
  }
}
// rays: n x 7 (origin, direction, time)
void orc_intersect(void* h, const float* rays, size_t n, float tmin, float tmax, uint8_t* hit, float* t, float* point, float* normal,
                   uint8_t* front_face, int32_t* mat, int32_t* leaf) {
  auto* s = static_cast<Scene*>(h);
  for (size_t i = 0; i < n; i++) {
    const float* r = rays + i * 7;
    Ray ray{V3{r[0], r[1], r[2]}, V3{r[3], r[4], r[5]}, r[6]};
    HitRecord rec;
    bool got = s->root->Hit(*s, ray, Interval{tmin, tmax}, rec);
    hit[i] = got;
    t[i] = got ? rec.t : 0;
    for (int k = 0; k < 3; k++) {
      point[i * 3 + k] = got ? rec.point[k] : 0;
      normal[i * 3 + k] = got ? rec.normal[k] : 0;
    }
    front_face[i] = got && rec.front_face;
    mat[i] = got ? rec.material : -1;
    if (leaf) leaf[i] = got ? rec.leaf_id : -1;
  }
}
// RayTracer::Update's per-pixel body (RayTracer.cpp:57-67) for frames [frame0, frame0+nframes), rows striped over threads.
void orc_render(void* h, int frame0, int nframes, int max_depth, int nthreads, double* sum, double* sumsq, uint64_t* n_rays, double* seconds) {
  auto* s = static_cast<Scene*>(h);
  const Camera& cam = s->cam;
  const int W = cam.w, H = cam.h, sq = cam.sq;
  if (nthreads < 1) nthreads = 1;
  std::vector<uint64_t> counts(nthreads, 0);
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> pool;
  for (int tid = 0; tid < nthreads; tid++) {
    pool.emplace_back([&, tid]() {
      tl_rays = 0;
      for (int f = frame0; f < frame0 + nframes; f++) {
        int s_i = f % sq, s_j = f / sq % sq;
        for (int y = tid; y < H; y += nthreads)
          for (int x = 0; x < W; x++) {
            V3 c = RayColor(cam.GetRay(x, y, s_i, s_j), max_depth, *s);
            size_t idx = (static_cast<size_t>(y) * W + x) * 3;
            for (int k = 0; k < 3; k++) {
              double v = c[k];
              sum[idx + k] += v;
              if (sumsq) sumsq[idx + k] += v * v;
            }
          }
      }
      counts[tid] = tl_rays;
    });
  }
  for (auto& th : pool) th.join();
  auto t1 = std::chrono::steady_clock::now();
  uint64_t total = 0;
  for (auto c : counts) total += c;
  if (n_rays) *n_rays = total;
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
