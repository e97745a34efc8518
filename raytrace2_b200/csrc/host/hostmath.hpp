// Host-side float vector / matrix helpers for the scene compiler.
//
// Everything the host precomputes for the device (quad plane constants, instance matrices, camera block, reference
// AABBs for the BVH-order replay) must round exactly like the reference, whose vector maths is GLM
// (third-party, unpinned: vcpkg.json:3-5).  The functions below therefore follow GLM's published operation order:
// dot = (x*x' + y*y') + z*z', normalize = v * (1 / sqrt(dot)), cross as written, mat4*vec4 =
// (m0*x + m1*y) + (m2*z + m3*w), mat4*mat4 left-associated column sums, inverse by cofactors, angleAxis / toMat4 as in
// glm/gtx/quaternion.  Compile with -ffp-contract=off.
#pragma once
#include <cmath>
#include <limits>

namespace rt2 {

struct V3 {
  float x{0}, y{0}, z{0};
  float& operator[](int i) { return (&x)[i]; }
  const float& operator[](int i) const { return (&x)[i]; }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float Dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline V3 Cross(V3 x, V3 y) { return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
inline V3 Normalize(V3 v) { return v * (1.0f / std::sqrt(Dot(v, v))); }

struct V4 {
  float x{0}, y{0}, z{0}, w{0};
  float& operator[](int i) { return (&x)[i]; }
  const float& operator[](int i) const { return (&x)[i]; }
};
inline V4 operator+(V4 a, V4 b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline V4 operator-(V4 a, V4 b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
inline V4 operator*(V4 a, V4 b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
inline V4 operator*(V4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }

// column-major: c[col][row], like glm::mat4
struct M4 {
  V4 c[4];
  V4& operator[](int i) { return c[i]; }
  const V4& operator[](int i) const { return c[i]; }
  static M4 Identity() {
    M4 m;
    m.c[0] = {1, 0, 0, 0};
    m.c[1] = {0, 1, 0, 0};
    m.c[2] = {0, 0, 1, 0};
    m.c[3] = {0, 0, 0, 1};
    return m;
  }
};

inline V4 Mul(const M4& m, V4 v) {
  V4 add0 = m[0] * v.x + m[1] * v.y;
  V4 add1 = m[2] * v.z + m[3] * v.w;
  return add0 + add1;
}
inline V3 MulPoint(const M4& m, V3 p) {
  V4 r = Mul(m, V4{p.x, p.y, p.z, 1.0f});
  return {r.x, r.y, r.z};
}
inline M4 Mul(const M4& a, const M4& b) {
  M4 r;
  for (int j = 0; j < 4; j++) r[j] = ((a[0] * b[j][0] + a[1] * b[j][1]) + a[2] * b[j][2]) + a[3] * b[j][3];
  return r;
}
inline M4 Translate(const M4& m, V3 v) {
  M4 r = m;
  r[3] = ((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3];
  return r;
}
inline M4 Scale(const M4& m, V3 v) {
  M4 r;
  r[0] = m[0] * v.x;
  r[1] = m[1] * v.y;
  r[2] = m[2] * v.z;
  r[3] = m[3];
  return r;
}
struct Quat {
  float w{1}, x{0}, y{0}, z{0};
};
inline float Radians(float deg) { return deg * static_cast<float>(0.01745329251994329576923690768489); }
inline Quat AngleAxis(float angle, V3 axis) {
  float s = std::sin(angle * 0.5f);
  return {std::cos(angle * 0.5f), axis.x * s, axis.y * s, axis.z * s};
}
inline M4 ToMat4(const Quat& q) {
  float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z, qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z;
  float qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
  M4 r = M4::Identity();
  r[0][0] = 1.0f - 2.0f * (qyy + qzz);
  r[0][1] = 2.0f * (qxy + qwz);
  r[0][2] = 2.0f * (qxz - qwy);
  r[1][0] = 2.0f * (qxy - qwz);
  r[1][1] = 1.0f - 2.0f * (qxx + qzz);
  r[1][2] = 2.0f * (qyz + qwx);
  r[2][0] = 2.0f * (qxz + qwy);
  r[2][1] = 2.0f * (qyz - qwx);
  r[2][2] = 1.0f - 2.0f * (qxx + qyy);
  return r;
}
inline M4 Inverse(const M4& m) {
  float c00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
  float c02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
  float c03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
  float c04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
  float c06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
  float c07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
  float c08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
  float c10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
  float c11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
  float c12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
  float c14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
  float c15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
  float c16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
  float c18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
  float c19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
  float c20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
  float c22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
  float c23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
  V4 f0{c00, c00, c02, c03}, f1{c04, c04, c06, c07}, f2{c08, c08, c10, c11};
  V4 f3{c12, c12, c14, c15}, f4{c16, c16, c18, c19}, f5{c20, c20, c22, c23};
  V4 v0{m[1][0], m[0][0], m[0][0], m[0][0]}, v1{m[1][1], m[0][1], m[0][1], m[0][1]};
  V4 v2{m[1][2], m[0][2], m[0][2], m[0][2]}, v3{m[1][3], m[0][3], m[0][3], m[0][3]};
  V4 i0 = (v1 * f0 - v2 * f1) + v3 * f2;
  V4 i1 = (v0 * f0 - v2 * f3) + v3 * f4;
  V4 i2 = (v0 * f1 - v1 * f3) + v3 * f5;
  V4 i3 = (v0 * f2 - v1 * f4) + v2 * f5;
  V4 sa{+1, -1, +1, -1}, sb{-1, +1, -1, +1};
  M4 inv;
  inv[0] = i0 * sa;
  inv[1] = i1 * sb;
  inv[2] = i2 * sa;
  inv[3] = i3 * sb;
  V4 row0{inv[0][0], inv[1][0], inv[2][0], inv[3][0]};
  V4 dot0 = m[0] * row0;
  float dot1 = (dot0.x + dot0.y) + (dot0.z + dot0.w);
  float one_over_det = 1.0f / dot1;
  M4 r;
  for (int i = 0; i < 4; i++) r[i] = inv[i] * one_over_det;
  return r;
}

constexpr float kInfinity = std::numeric_limits<float>::max();  // Defs.hpp:17 — FLT_MAX, not IEEE inf

// Interval / AABB exactly as the reference builds them (Interval.hpp:6-24, AABB.hpp:9-65); needed bit-for-bit to
// replay the reference's BVH construction order (BVH.cpp:10-38) and derive the span-1 ("Q2") flags.
struct RefInterval {
  float min{kInfinity}, max{-kInfinity};
  float Size() const { return max - min; }
  RefInterval Expand(float delta) const {
    float padding = delta / 2.0f;
    return {min - padding, max + padding};
  }
  static RefInterval Union(const RefInterval& a, const RefInterval& b) { return {std::fmin(a.min, b.min), std::fmax(a.max, b.max)}; }
};
struct RefAABB {
  RefInterval x, y, z;
  RefAABB() = default;
  RefAABB(V3 a, V3 b)
      : x{std::fmin(a.x, b.x), std::fmax(a.x, b.x)}, y{std::fmin(a.y, b.y), std::fmax(a.y, b.y)}, z{std::fmin(a.z, b.z), std::fmax(a.z, b.z)} {
    Pad();
  }
  RefAABB(const RefAABB& a, const RefAABB& b)
      : x(RefInterval::Union(a.x, b.x)), y(RefInterval::Union(a.y, b.y)), z(RefInterval::Union(a.z, b.z)) {
    Pad();
  }
  const RefInterval& Axis(int n) const { return n == 0 ? x : (n == 1 ? y : z); }
  V3 Min() const { return {x.min, y.min, z.min}; }
  V3 Max() const { return {x.max, y.max, z.max}; }
  int LongestAxis() const {
    if (x.Size() > y.Size()) return x.Size() > z.Size() ? 0 : 2;
    return y.Size() > z.Size() ? 1 : 2;
  }
  void Pad() {
    constexpr float kDelta = 0.0001f;
    if (x.Size() < kDelta) x = x.Expand(kDelta);
    if (y.Size() < kDelta) y = y.Expand(kDelta);
    if (z.Size() < kDelta) z = z.Expand(kDelta);
  }
};

}  // namespace rt2
