// oracle/shim/glm/glm_subset.hpp — TEST INFRASTRUCTURE (never linked into the product).
//
// A hand-written subset of the GLM maths library (third-party dependency of the reference,
// vcpkg port "glm", unpinned: /root/reference/vcpkg.json:3-5), covering exactly the types and
// functions the reference's hot path uses (src/cpu_raytrace/*, src/Serialize.cpp, src/Util.cpp).
// GLM itself is not in this image and cannot be fetched, so the oracle build compiles the
// reference's own sources against this header.  Every function follows GLM's published operation
// order (dot = (x*x' + y*y') + z*z', normalize = v * (1/sqrt(dot)), mat4*vec4 =
// (m0*x + m1*y) + (m2*z + m3*w), inverse(mat4) by cofactors, ...) so that rounding matches a build
// against real GLM compiled with -ffp-contract=off.  Parity at this boundary is UNPINNED by the
// reference (it has no tests): this shim *defines* the oracle's rounding.
#pragma once
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>

namespace glm {

template <typename T> struct tvec2;
template <typename T> struct tvec3;
template <typename T> struct tvec4;

template <typename T>
struct tvec2 {
  T x{}, y{};
  constexpr tvec2() = default;
  template <typename A> constexpr explicit tvec2(A s) : x(static_cast<T>(s)), y(static_cast<T>(s)) {}
  template <typename A, typename B>
  constexpr tvec2(A a, B b) : x(static_cast<T>(a)), y(static_cast<T>(b)) {}
  template <typename U> constexpr tvec2(const tvec2<U>& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)) {}
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
};

template <typename T>
struct tvec3 {
  T x{}, y{}, z{};
  constexpr tvec3() = default;
  template <typename A>
  constexpr tvec3(A s) : x(static_cast<T>(s)), y(static_cast<T>(s)), z(static_cast<T>(s)) {}
  template <typename A, typename B, typename C>
  constexpr tvec3(A a, B b, C c) : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)) {}
  template <typename U>
  constexpr tvec3(const tvec3<U>& v) : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}
  template <typename U>
  constexpr explicit tvec3(const tvec4<U>& v);
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
  tvec3& operator+=(const tvec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
  tvec3& operator-=(const tvec3& o) { x -= o.x; y -= o.y; z -= o.z; return *this; }
  template <typename S> tvec3& operator*=(S s) { x *= static_cast<T>(s); y *= static_cast<T>(s); z *= static_cast<T>(s); return *this; }
  tvec3& operator*=(const tvec3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
  template <typename S> tvec3& operator/=(S s) { x /= static_cast<T>(s); y /= static_cast<T>(s); z /= static_cast<T>(s); return *this; }
};

template <typename T>
struct tvec4 {
  T x{}, y{}, z{}, w{};
  constexpr tvec4() = default;
  template <typename A>
  constexpr explicit tvec4(A s) : x(static_cast<T>(s)), y(static_cast<T>(s)), z(static_cast<T>(s)), w(static_cast<T>(s)) {}
  template <typename A, typename B, typename C, typename D>
  constexpr tvec4(A a, B b, C c, D d)
      : x(static_cast<T>(a)), y(static_cast<T>(b)), z(static_cast<T>(c)), w(static_cast<T>(d)) {}
  template <typename U, typename W>
  constexpr tvec4(const tvec3<U>& v, W ww)
      : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)), w(static_cast<T>(ww)) {}
  T& operator[](int i) { return (&x)[i]; }
  const T& operator[](int i) const { return (&x)[i]; }
};

template <typename T>
template <typename U>
constexpr tvec3<T>::tvec3(const tvec4<U>& v)
    : x(static_cast<T>(v.x)), y(static_cast<T>(v.y)), z(static_cast<T>(v.z)) {}

// ---- vec3 arithmetic (component-wise, like GLM) ----
template <typename T> constexpr tvec3<T> operator+(const tvec3<T>& a, const tvec3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T> constexpr tvec3<T> operator-(const tvec3<T>& a, const tvec3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T> constexpr tvec3<T> operator*(const tvec3<T>& a, const tvec3<T>& b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
template <typename T> constexpr tvec3<T> operator/(const tvec3<T>& a, const tvec3<T>& b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
template <typename T> constexpr tvec3<T> operator-(const tvec3<T>& a) { return {-a.x, -a.y, -a.z}; }
// scalar forms: GLM converts the scalar to T first (vec<3,T> op T)
template <typename T> constexpr tvec3<T> operator*(const tvec3<T>& a, T s) { return {a.x * s, a.y * s, a.z * s}; }
template <typename T> constexpr tvec3<T> operator*(T s, const tvec3<T>& a) { return {s * a.x, s * a.y, s * a.z}; }
template <typename T> constexpr tvec3<T> operator/(const tvec3<T>& a, T s) { return {a.x / s, a.y / s, a.z / s}; }
template <typename T> constexpr tvec3<T> operator+(const tvec3<T>& a, T s) { return {a.x + s, a.y + s, a.z + s}; }
template <typename T> constexpr tvec3<T> operator-(const tvec3<T>& a, T s) { return {a.x - s, a.y - s, a.z - s}; }
// mixed scalar types (int / double literals against float vectors): convert to T, as GLM's
// templated scalar operators do.
#define GLM_SUBSET_MIXED(S)                                                                                        \
  template <typename T> constexpr tvec3<T> operator*(const tvec3<T>& a, S s) { return a * static_cast<T>(s); }     \
  template <typename T> constexpr tvec3<T> operator*(S s, const tvec3<T>& a) { return static_cast<T>(s) * a; }     \
  template <typename T> constexpr tvec3<T> operator/(const tvec3<T>& a, S s) { return a / static_cast<T>(s); }
GLM_SUBSET_MIXED(int)
#undef GLM_SUBSET_MIXED
inline tvec3<float> operator*(const tvec3<float>& a, double s) { return a * static_cast<float>(s); }
inline tvec3<float> operator*(double s, const tvec3<float>& a) { return static_cast<float>(s) * a; }
inline tvec3<float> operator/(const tvec3<float>& a, double s) { return a / static_cast<float>(s); }

template <typename T> constexpr tvec4<T> operator+(const tvec4<T>& a, const tvec4<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
template <typename T> constexpr tvec4<T> operator-(const tvec4<T>& a, const tvec4<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
template <typename T> constexpr tvec4<T> operator*(const tvec4<T>& a, const tvec4<T>& b) { return {a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w}; }
template <typename T> constexpr tvec4<T> operator*(const tvec4<T>& a, T s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }

template <typename T> constexpr tvec2<T> operator+(const tvec2<T>& a, const tvec2<T>& b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> constexpr tvec2<T> operator-(const tvec2<T>& a, const tvec2<T>& b) { return {a.x - b.x, a.y - b.y}; }
template <typename T> constexpr tvec2<T> operator*(const tvec2<T>& a, T s) { return {a.x * s, a.y * s}; }
template <typename T> constexpr bool operator==(const tvec2<T>& a, const tvec2<T>& b) { return a.x == b.x && a.y == b.y; }
template <typename T> constexpr bool operator!=(const tvec2<T>& a, const tvec2<T>& b) { return !(a == b); }

using vec2 = tvec2<float>;
using vec3 = tvec3<float>;
using vec4 = tvec4<float>;
using dvec2 = tvec2<double>;
using dvec3 = tvec3<double>;
using dvec4 = tvec4<double>;
using ivec2 = tvec2<int>;
using ivec3 = tvec3<int>;
using u8vec4 = tvec4<std::uint8_t>;

// ---- scalar functions ----
using std::pow;  // GLM pulls std::pow into its namespace, so glm::pow(float, int) is std::pow -> double
template <typename T> constexpr T min(T a, T b) { return (b < a) ? b : a; }
template <typename T> constexpr T max(T a, T b) { return (a < b) ? b : a; }
template <typename T> constexpr T clamp(T x, T lo, T hi) { return min(max(x, lo), hi); }
inline float sqrt(float x) { return std::sqrt(x); }
inline double sqrt(double x) { return std::sqrt(x); }
inline float tan(float x) { return std::tan(x); }
inline double tan(double x) { return std::tan(x); }
inline float acos(float x) { return std::acos(x); }
inline double acos(double x) { return std::acos(x); }
inline float floor(float x) { return std::floor(x); }
inline float radians(float deg) { return deg * static_cast<float>(0.01745329251994329576923690768489); }
inline double radians(double deg) { return deg * 0.01745329251994329576923690768489; }

// ---- vector functions ----
template <typename T> constexpr T dot(const tvec3<T>& a, const tvec3<T>& b) {
  tvec3<T> tmp(a * b);
  return tmp.x + tmp.y + tmp.z;
}
template <typename T> constexpr T dot(const tvec4<T>& a, const tvec4<T>& b) {
  tvec4<T> tmp(a * b);
  return (tmp.x + tmp.y) + (tmp.z + tmp.w);
}
template <typename T> inline T length(const tvec3<T>& v) { return std::sqrt(dot(v, v)); }
template <typename T> inline tvec3<T> normalize(const tvec3<T>& v) {
  return v * (static_cast<T>(1) / std::sqrt(dot(v, v)));
}
template <typename T> constexpr tvec3<T> cross(const tvec3<T>& x, const tvec3<T>& y) {
  return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y};
}
template <typename T> inline tvec3<T> floor(const tvec3<T>& v) { return {std::floor(v.x), std::floor(v.y), std::floor(v.z)}; }
template <typename T> constexpr tvec3<T> clamp(const tvec3<T>& v, T lo, T hi) {
  return {clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)};
}

// ---- matrices (column-major, m[col][row]) ----
template <typename T>
struct tmat3 {
  tvec3<T> c[3];
  constexpr tmat3() : tmat3(static_cast<T>(1)) {}
  constexpr explicit tmat3(T d) : c{{d, 0, 0}, {0, d, 0}, {0, 0, d}} {}
  constexpr tmat3(const tvec3<T>& a, const tvec3<T>& b, const tvec3<T>& cc) : c{a, b, cc} {}
  tvec3<T>& operator[](int i) { return c[i]; }
  const tvec3<T>& operator[](int i) const { return c[i]; }
};

template <typename T>
struct tmat4 {
  tvec4<T> c[4];
  constexpr tmat4() : tmat4(static_cast<T>(1)) {}
  template <typename A>
  constexpr explicit tmat4(A dd)
      : c{{static_cast<T>(dd), 0, 0, 0}, {0, static_cast<T>(dd), 0, 0}, {0, 0, static_cast<T>(dd), 0}, {0, 0, 0, static_cast<T>(dd)}} {}
  constexpr tmat4(const tvec4<T>& a, const tvec4<T>& b, const tvec4<T>& cc, const tvec4<T>& d) : c{a, b, cc, d} {}
  constexpr explicit tmat4(const tmat3<T>& m)
      : c{{m[0].x, m[0].y, m[0].z, 0}, {m[1].x, m[1].y, m[1].z, 0}, {m[2].x, m[2].y, m[2].z, 0}, {0, 0, 0, 1}} {}
  tvec4<T>& operator[](int i) { return c[i]; }
  const tvec4<T>& operator[](int i) const { return c[i]; }
};

// mat3 from mat4 = upper-left block.  Implicit: the reference assigns transpose(inverse(mat4)) to a mat3.
template <typename T>
struct tmat3_from4 : tmat3<T> {};
using mat3_base = tmat3<float>;
struct mat3 : tmat3<float> {
  using tmat3<float>::tmat3;
  constexpr mat3() = default;
  constexpr mat3(const tmat3<float>& m) : tmat3<float>(m) {}
  constexpr mat3(const tmat4<float>& m)
      : tmat3<float>(tvec3<float>(m[0]), tvec3<float>(m[1]), tvec3<float>(m[2])) {}
};
using mat4 = tmat4<float>;
using dmat4 = tmat4<double>;
using dmat3 = tmat3<double>;

template <typename T> constexpr tvec4<T> operator*(const tmat4<T>& m, const tvec4<T>& v) {
  const tvec4<T> Mov0(v[0]), Mov1(v[1]);
  const tvec4<T> Mul0 = m[0] * Mov0, Mul1 = m[1] * Mov1;
  const tvec4<T> Add0 = Mul0 + Mul1;
  const tvec4<T> Mov2(v[2]), Mov3(v[3]);
  const tvec4<T> Mul2 = m[2] * Mov2, Mul3 = m[3] * Mov3;
  const tvec4<T> Add1 = Mul2 + Mul3;
  return Add0 + Add1;
}
template <typename T> constexpr tvec3<T> operator*(const tmat3<T>& m, const tvec3<T>& v) {
  return {m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z, m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
          m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z};
}
template <typename T> constexpr tmat4<T> operator*(const tmat4<T>& m1, const tmat4<T>& m2) {
  const tvec4<T> A0 = m1[0], A1 = m1[1], A2 = m1[2], A3 = m1[3];
  tmat4<T> r;
  for (int j = 0; j < 4; j++) {
    const tvec4<T> B = m2[j];
    r[j] = A0 * B[0] + A1 * B[1] + A2 * B[2] + A3 * B[3];
  }
  return r;
}
template <typename T> constexpr tmat4<T> operator*(const tmat4<T>& m, T s) { return {m[0] * s, m[1] * s, m[2] * s, m[3] * s}; }

template <typename T> constexpr tmat4<T> transpose(const tmat4<T>& m) {
  tmat4<T> r;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) r[i][j] = m[j][i];
  return r;
}

template <typename T> constexpr tmat4<T> inverse(const tmat4<T>& m) {
  T Coef00 = m[2][2] * m[3][3] - m[3][2] * m[2][3];
  T Coef02 = m[1][2] * m[3][3] - m[3][2] * m[1][3];
  T Coef03 = m[1][2] * m[2][3] - m[2][2] * m[1][3];
  T Coef04 = m[2][1] * m[3][3] - m[3][1] * m[2][3];
  T Coef06 = m[1][1] * m[3][3] - m[3][1] * m[1][3];
  T Coef07 = m[1][1] * m[2][3] - m[2][1] * m[1][3];
  T Coef08 = m[2][1] * m[3][2] - m[3][1] * m[2][2];
  T Coef10 = m[1][1] * m[3][2] - m[3][1] * m[1][2];
  T Coef11 = m[1][1] * m[2][2] - m[2][1] * m[1][2];
  T Coef12 = m[2][0] * m[3][3] - m[3][0] * m[2][3];
  T Coef14 = m[1][0] * m[3][3] - m[3][0] * m[1][3];
  T Coef15 = m[1][0] * m[2][3] - m[2][0] * m[1][3];
  T Coef16 = m[2][0] * m[3][2] - m[3][0] * m[2][2];
  T Coef18 = m[1][0] * m[3][2] - m[3][0] * m[1][2];
  T Coef19 = m[1][0] * m[2][2] - m[2][0] * m[1][2];
  T Coef20 = m[2][0] * m[3][1] - m[3][0] * m[2][1];
  T Coef22 = m[1][0] * m[3][1] - m[3][0] * m[1][1];
  T Coef23 = m[1][0] * m[2][1] - m[2][0] * m[1][1];
  tvec4<T> Fac0(Coef00, Coef00, Coef02, Coef03), Fac1(Coef04, Coef04, Coef06, Coef07);
  tvec4<T> Fac2(Coef08, Coef08, Coef10, Coef11), Fac3(Coef12, Coef12, Coef14, Coef15);
  tvec4<T> Fac4(Coef16, Coef16, Coef18, Coef19), Fac5(Coef20, Coef20, Coef22, Coef23);
  tvec4<T> Vec0(m[1][0], m[0][0], m[0][0], m[0][0]), Vec1(m[1][1], m[0][1], m[0][1], m[0][1]);
  tvec4<T> Vec2(m[1][2], m[0][2], m[0][2], m[0][2]), Vec3(m[1][3], m[0][3], m[0][3], m[0][3]);
  tvec4<T> Inv0(Vec1 * Fac0 - Vec2 * Fac1 + Vec3 * Fac2);
  tvec4<T> Inv1(Vec0 * Fac0 - Vec2 * Fac3 + Vec3 * Fac4);
  tvec4<T> Inv2(Vec0 * Fac1 - Vec1 * Fac3 + Vec3 * Fac5);
  tvec4<T> Inv3(Vec0 * Fac2 - Vec1 * Fac4 + Vec2 * Fac5);
  tvec4<T> SignA(+1, -1, +1, -1), SignB(-1, +1, -1, +1);
  tmat4<T> Inverse(Inv0 * SignA, Inv1 * SignB, Inv2 * SignA, Inv3 * SignB);
  tvec4<T> Row0(Inverse[0][0], Inverse[1][0], Inverse[2][0], Inverse[3][0]);
  tvec4<T> Dot0(m[0] * Row0);
  T Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
  T OneOverDeterminant = static_cast<T>(1) / Dot1;
  return Inverse * OneOverDeterminant;
}

template <typename T> constexpr tmat4<T> translate(const tmat4<T>& m, const tvec3<T>& v) {
  tmat4<T> r(m);
  r[3] = m[0] * v[0] + m[1] * v[1] + m[2] * v[2] + m[3];
  return r;
}
template <typename T> constexpr tmat4<T> scale(const tmat4<T>& m, const tvec3<T>& v) {
  return {m[0] * v[0], m[1] * v[1], m[2] * v[2], m[3]};
}

// ---- quaternion (w, x, y, z) ----
template <typename T>
struct tquat {
  // GLM leaves a default-constructed quat uninitialised, and the reference reads it when a transform has no
  // "rotation" key (Serialize.cpp:114,125) — undefined behaviour whose evident intent is "no rotation".  The oracle
  // pins that intent: identity.
  T x{0}, y{0}, z{0}, w{1};
  constexpr tquat() = default;
  constexpr tquat(T ww, T xx, T yy, T zz) : x(xx), y(yy), z(zz), w(ww) {}
  constexpr tquat(T ww, const tvec3<T>& v) : x(v.x), y(v.y), z(v.z), w(ww) {}
};
using quat = tquat<float>;
using dquat = tquat<double>;

template <typename T> inline tquat<T> angleAxis(T angle, const tvec3<T>& v) {
  const T a(angle);
  const T s = std::sin(a * static_cast<T>(0.5));
  return tquat<T>(std::cos(a * static_cast<T>(0.5)), v * s);
}
template <typename T> constexpr tmat4<T> toMat4(const tquat<T>& q) {
  tmat3<T> R(static_cast<T>(1));
  T qxx(q.x * q.x), qyy(q.y * q.y), qzz(q.z * q.z), qxz(q.x * q.z), qxy(q.x * q.y), qyz(q.y * q.z);
  T qwx(q.w * q.x), qwy(q.w * q.y), qwz(q.w * q.z);
  R[0][0] = T(1) - T(2) * (qyy + qzz);
  R[0][1] = T(2) * (qxy + qwz);
  R[0][2] = T(2) * (qxz - qwy);
  R[1][0] = T(2) * (qxy - qwz);
  R[1][1] = T(1) - T(2) * (qxx + qzz);
  R[1][2] = T(2) * (qyz + qwx);
  R[2][0] = T(2) * (qxz + qwy);
  R[2][1] = T(2) * (qyz - qwx);
  R[2][2] = T(1) - T(2) * (qxx + qyy);
  return tmat4<T>(R);
}

}  // namespace glm
