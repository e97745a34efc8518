// Device-side arithmetic policies, Philox RNG and small vector helpers.
//
// ExactMath reproduces the reference's IEEE binary32 arithmetic operation by operation (separate round-to-nearest
// mul / add / div / sqrt, no FMA contraction) so that sphere / quad / instance Hit() results are bit-identical with the
// reference compiled with -ffp-contract=off.  FastMath lets ptxas contract to FMA.  The operation ORDER
// follows the reference sources and GLM (dot = (x*x' + y*y') + z*z', normalize = v * (1/sqrt(dot)), ...).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rt2dev {

struct ExactMath {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
};
struct FastMath {
  static __device__ __forceinline__ float mul(float a, float b) { return a * b; }
  static __device__ __forceinline__ float add(float a, float b) { return a + b; }
  static __device__ __forceinline__ float sub(float a, float b) { return a - b; }
  static __device__ __forceinline__ float div(float a, float b) { return __fdividef(a, b); }
  static __device__ __forceinline__ float sqrt(float a) { return __fsqrt_rn(a); }
};

struct F3 {
  float x, y, z;
};
__device__ __forceinline__ F3 make_f3(float x, float y, float z) { return F3{x, y, z}; }
__device__ __forceinline__ F3 make_f3(float4 v) { return F3{v.x, v.y, v.z}; }

template <class M> __device__ __forceinline__ F3 vadd(F3 a, F3 b) { return {M::add(a.x, b.x), M::add(a.y, b.y), M::add(a.z, b.z)}; }
template <class M> __device__ __forceinline__ F3 vsub(F3 a, F3 b) { return {M::sub(a.x, b.x), M::sub(a.y, b.y), M::sub(a.z, b.z)}; }
template <class M> __device__ __forceinline__ F3 vscale(F3 a, float s) { return {M::mul(a.x, s), M::mul(a.y, s), M::mul(a.z, s)}; }
template <class M> __device__ __forceinline__ F3 vdivs(F3 a, float s) { return {M::div(a.x, s), M::div(a.y, s), M::div(a.z, s)}; }
__device__ __forceinline__ F3 vneg(F3 a) { return {-a.x, -a.y, -a.z}; }
// glm::dot(vec3): (x*x' + y*y') + z*z'
template <class M> __device__ __forceinline__ float vdot(F3 a, F3 b) {
  return M::add(M::add(M::mul(a.x, b.x), M::mul(a.y, b.y)), M::mul(a.z, b.z));
}
// glm::cross
template <class M> __device__ __forceinline__ F3 vcross(F3 x, F3 y) {
  return {M::sub(M::mul(x.y, y.z), M::mul(y.y, x.z)), M::sub(M::mul(x.z, y.x), M::mul(y.z, x.x)),
          M::sub(M::mul(x.x, y.y), M::mul(y.x, x.y))};
}
// glm::normalize: v * (1 / sqrt(dot(v, v)))
template <class M> __device__ __forceinline__ F3 vnormalize(F3 v) {
  float s = M::div(1.0f, M::sqrt(vdot<M>(v, v)));
  return vscale<M>(v, s);
}
// Ray::At (Ray.hpp:7): origin + direction * t
template <class M> __device__ __forceinline__ F3 ray_at(F3 o, F3 d, float t) { return vadd<M>(o, vscale<M>(d, t)); }

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based: key = (pixel, seed), counter = (frame, bounce|stream, ...) ----
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
  const uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
  // two rounds per trip: the rounds are one dependent chain, so full unrolling buys no ILP — it only cost 1.2 KB of code per
  // call site in kernels that are instruction-cache bound (k_finish_shade)
#pragma unroll 2
  for (int r = 0; r < 10; r++) {
    uint32_t hi0 = __umulhi(kM0, c.x), lo0 = kM0 * c.x;
    uint32_t hi1 = __umulhi(kM1, c.z), lo1 = kM1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += kW0;
    k.y += kW1;
  }
  return c;
}
// U[0,1) with 24 bits, like uniform_real_distribution<float> over minstd (Math.hpp:9-13)
__device__ __forceinline__ float u01(uint32_t x) { return static_cast<float>(x >> 8) * (1.0f / 16777216.0f); }

enum : uint32_t { kStreamCamera = 0, kStreamLens = 1, kStreamScatter = 2, kStreamMedium = 8 };

struct RngKey {
  uint32_t pixel;  // key.x
  uint32_t frame;  // counter.x
  uint32_t seed_lo, seed_hi;
};
__device__ __forceinline__ uint4 rng_draw(const RngKey& k, uint32_t bounce, uint32_t stream) {
  return philox4x32_10(make_uint4(k.frame, bounce | (stream << 16), k.seed_hi, 0x52543242u), make_uint2(k.pixel, k.seed_lo));
}

// math::RandUnitVec3 (Math.hpp:43) = normalised uniform-in-ball rejection sample = uniform direction; sampled directly.
__device__ __forceinline__ F3 unit_vector(float u1, float u2) {
  float z = 1.0f - 2.0f * u1;
  float r = sqrtf(fmaxf(0.0f, 1.0f - z * z));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return {r * c, r * s, z};
}

}  // namespace rt2dev
