// Materials, textures and Perlin noise on the device: replaces Material<T>::Scatter/Emit (src/cpu_raytrace/Material.cpp),
// texture::*::Value (Texture.cpp) and PerlinNoiseGen (PerlinNoiseGen.cpp).  Shading consumes random numbers, so parity
// with the reference is statistical; plain (FMA-contracted) float arithmetic is used here.
#pragma once
#include "rt_trace.cuh"

namespace rt2dev {

// PerlinNoiseGen::Noise + PerlinInterp (PerlinNoiseGen.cpp:10-26,66-88), operation by operation in the reference's order and
// without FMA contraction (ExactMath), so that with the same tables the value is the reference's bit for bit — which is what
// lets rt2_texture_value be checked at fixed points (the marble's sin() then sees the same argument; sinf vs libm is the
// only difference left).  The weights i*uu + (1-i)*(1-uu) reduce exactly to uu / 1-uu.
__device__ __forceinline__ float perlin_noise(const rt2_perlin* __restrict__ P, F3 p) {
  using E = ExactMath;
  const float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
  const float u = E::sub(p.x, fx), v = E::sub(p.y, fy), w = E::sub(p.z, fz);
  const int i = static_cast<int>(fx), j = static_cast<int>(fy), k = static_cast<int>(fz);
  const float uu = E::mul(E::mul(u, u), E::sub(3.0f, E::mul(2.0f, u)));
  const float vv = E::mul(E::mul(v, v), E::sub(3.0f, E::mul(2.0f, v)));
  const float ww = E::mul(E::mul(w, w), E::sub(3.0f, E::mul(2.0f, w)));
  float accum = 0.0f;
#pragma unroll
  for (int di = 0; di < 2; di++) {
    const int px = __ldg(&P->perm_x[(i + di) & 255]);
#pragma unroll
    for (int dj = 0; dj < 2; dj++) {
      const int py = __ldg(&P->perm_y[(j + dj) & 255]);
#pragma unroll
      for (int dk = 0; dk < 2; dk++) {
        const int pz = __ldg(&P->perm_z[(k + dk) & 255]);
        const float4 c = __ldg(reinterpret_cast<const float4*>(&P->vec[(px ^ py ^ pz) & 255][0]));
        const float wx = E::sub(u, static_cast<float>(di)), wy = E::sub(v, static_cast<float>(dj)), wz = E::sub(w, static_cast<float>(dk));
        const float a = di ? uu : E::sub(1.0f, uu), b = dj ? vv : E::sub(1.0f, vv), cc = dk ? ww : E::sub(1.0f, ww);
        const float dot = E::add(E::add(E::mul(c.x, wx), E::mul(c.y, wy)), E::mul(c.z, wz));  // glm::dot
        accum = E::add(accum, E::mul(E::mul(E::mul(a, b), cc), dot));
      }
    }
  }
  return accum;
}

// PerlinNoiseGen::Turb(p, 7) (PerlinNoiseGen.cpp:52-64)
__device__ __forceinline__ float perlin_turb(const rt2_perlin* __restrict__ P, F3 p) {
  float accum = 0.0f, weight = 1.0f;
  for (int i = 0; i < 7; i++) {
    accum = ExactMath::add(accum, ExactMath::mul(weight, perlin_noise(P, p)));
    weight *= 0.5f;
    p = {p.x * 2.0f, p.y * 2.0f, p.z * 2.0f};
  }
  return fabsf(accum);
}

// Image texture lookup (schema extension; the book's image_texture::value): u clamped to [0,1], v flipped and clamped,
// nearest texel, texels already linear.
__device__ __forceinline__ F3 image_value(const DeviceScene& S, uint32_t image_idx, float u, float v) {
  const uint4 im = __ldg(S.images + image_idx);
  u = fminf(fmaxf(u, 0.0f), 1.0f);
  v = 1.0f - fminf(fmaxf(v, 0.0f), 1.0f);
  const uint32_t i = min(static_cast<uint32_t>(u * static_cast<float>(im.y)), im.y - 1u);
  const uint32_t j = min(static_cast<uint32_t>(v * static_cast<float>(im.z)), im.z - 1u);
  const float4 c = __ldg(S.image_texels + im.x + static_cast<size_t>(j) * im.y + i);
  return {c.x, c.y, c.z};
}
// (u, v) of a hit record travel through the queues as two 16-bit unorms
__device__ __forceinline__ uint32_t pack_uv16(float u, float v) {
  const uint32_t a = static_cast<uint32_t>(fminf(fmaxf(u, 0.0f), 1.0f) * 65535.0f + 0.5f);
  const uint32_t b = static_cast<uint32_t>(fminf(fmaxf(v, 0.0f), 1.0f) * 65535.0f + 0.5f);
  return a | (b << 16);
}
__device__ __forceinline__ float2 unpack_uv16(uint32_t w) {
  return make_float2(static_cast<float>(w & 0xFFFFu) * (1.0f / 65535.0f), static_cast<float>(w >> 16) * (1.0f / 65535.0f));
}

// texture::{SolidColor,Checker,Noise}::Value (Texture.hpp:14-17, Texture.cpp:7-22).  Checker recursion through texture
// indices is followed for at most 8 levels (the reference would recurse forever on a cycle).
__device__ __forceinline__ F3 texture_value(const DeviceScene& S, uint32_t tex_idx, F3 p, float u = 0.0f, float v = 0.0f) {
  for (int depth = 0; depth < 8; depth++) {
    const float4 t0 = __ldg(S.textures + 3 * tex_idx), t1 = __ldg(S.textures + 3 * tex_idx + 1);
    const uint32_t type = __float_as_uint(t0.x);
    if (type == RT2_TEX_CHECKER) {
      // glm::ivec3 i = glm::floor(inv_scale * p); (i.x + i.y + i.z) % 2 == 0 ? even : odd
      int ix = static_cast<int>(floorf(t1.w * p.x)), iy = static_cast<int>(floorf(t1.w * p.y)), iz = static_cast<int>(floorf(t1.w * p.z));
      tex_idx = ((ix + iy + iz) % 2 == 0) ? __float_as_uint(t0.y) : __float_as_uint(t0.z);
      continue;
    }
    if (type == RT2_TEX_IMAGE) return image_value(S, __float_as_uint(__ldg(S.textures + 3 * tex_idx + 2).y), u, v);
    if (type == RT2_TEX_NOISE) {
      const float4 t2 = __ldg(S.textures + 3 * tex_idx + 2);
      const rt2_perlin* P = S.perlin + __float_as_uint(t0.w);
      float f;
      // Texture.cpp:16-22: albedo * 0.5 * (1 + noise(scale * p))  |  albedo * 0.5 * (1 + sin(scale * p.z + 10 * turb(p)))
      if (__float_as_uint(t2.x) == 0u) {  // NoiseType::kPerlin
        f = ExactMath::add(1.0f, perlin_noise(P, F3{ExactMath::mul(t1.w, p.x), ExactMath::mul(t1.w, p.y), ExactMath::mul(t1.w, p.z)}));
      } else {  // kMarble
        f = ExactMath::add(1.0f, sinf(ExactMath::add(ExactMath::mul(t1.w, p.z), ExactMath::mul(10.0f, perlin_turb(P, p)))));
      }
      return {ExactMath::mul(ExactMath::mul(t1.x, 0.5f), f), ExactMath::mul(ExactMath::mul(t1.y, 0.5f), f), ExactMath::mul(ExactMath::mul(t1.z, 0.5f), f)};
    }
    return {t1.x, t1.y, t1.z};  // solid colour
  }
  return {0.0f, 0.0f, 0.0f};
}

// texture_value for textures whose chain holds no noise texture (solid colours, checkers of solid colours): the cheap
// subset the fused finish+shade kernel evaluates inline.  The host marks every material that can reach a noise texture
// as deferred (Renderer::UploadScene), so the noise branch is never needed here.
template <bool kImages = true>
__device__ __forceinline__ F3 texture_value_simple(const DeviceScene& S, uint32_t tex_idx, F3 p, float u, float v) {
#pragma unroll 1
  for (int depth = 0; depth < 8; depth++) {
    const float4 t0 = __ldg(S.textures + 3 * tex_idx), t1 = __ldg(S.textures + 3 * tex_idx + 1);
    if (kImages && __float_as_uint(t0.x) == RT2_TEX_IMAGE) return image_value(S, __float_as_uint(__ldg(S.textures + 3 * tex_idx + 2).y), u, v);
    if (__float_as_uint(t0.x) != RT2_TEX_CHECKER) return {t1.x, t1.y, t1.z};
    int ix = static_cast<int>(floorf(t1.w * p.x)), iy = static_cast<int>(floorf(t1.w * p.y)), iz = static_cast<int>(floorf(t1.w * p.z));
    tex_idx = ((ix + iy + iz) % 2 == 0) ? __float_as_uint(t0.y) : __float_as_uint(t0.z);
  }
  return {0.0f, 0.0f, 0.0f};
}

// math::NearZero (Math.hpp:61-64): |v_i| < 1e-8 (double literal)  <=>  |v_i| <= float(1e-8)
__device__ __forceinline__ bool near_zero(F3 v) {
  const float e = 9.99999993922529e-09f;
  return fabsf(v.x) <= e && fabsf(v.y) <= e && fabsf(v.z) <= e;
}
__device__ __forceinline__ float dot3(F3 a, F3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ F3 normalize3(F3 v) {
  float s = 1.0f / sqrtf(dot3(v, v));
  return {v.x * s, v.y * s, v.z * s};
}
// math::Reflect (Math.hpp:66)
__device__ __forceinline__ F3 reflect3(F3 v, F3 n) {
  float k = 2.0f * dot3(v, n);
  return {v.x - k * n.x, v.y - k * n.y, v.z - k * n.z};
}
// math::Refract (Math.hpp:68-73)
__device__ __forceinline__ F3 refract3(F3 uv, F3 n, float eta) {
  float cos_theta = fminf(dot3(vneg(uv), n), 1.0f);
  F3 perp = {eta * (uv.x + cos_theta * n.x), eta * (uv.y + cos_theta * n.y), eta * (uv.z + cos_theta * n.z)};
  float k = -sqrtf(fabsf(1.0f - dot3(perp, perp)));
  return {perp.x + k * n.x, perp.y + k * n.y, perp.z + k * n.z};
}

// Material<T>::Scatter for the scattering types.  Returns the attenuation; `dir` is the new direction (origin = hit
// point, time unchanged).  r = the bounce's scatter draw (x,y: unit vector, z: dielectric xi).
template <int kType>
__device__ __forceinline__ F3 scatter(const DeviceScene& S, const float4 m0, const float4 m1, F3 d_in, F3 p, F3 n, bool front_face,
                                      const uint4 r, F3& dir, float tu = 0.0f, float tv = 0.0f) {
  if (kType == RT2_MAT_LAMBERTIAN || kType == RT2_MAT_TEXTURE) {
    // Material.cpp:47-69
    F3 u = unit_vector(u01(r.x), u01(r.y));
    dir = {n.x + u.x, n.y + u.y, n.z + u.z};
    if (near_zero(dir)) dir = n;
    if (kType == RT2_MAT_LAMBERTIAN) return {m1.x, m1.y, m1.z};
    return texture_value(S, __float_as_uint(m0.y), p, tu, tv);
  }
  if (kType == RT2_MAT_METAL) {
    // Material.cpp:10-17: always scatters (no dot(scattered, normal) > 0 test)
    F3 u = unit_vector(u01(r.x), u01(r.y));
    F3 refl = normalize3(reflect3(d_in, n));
    dir = {refl.x + m0.z * u.x, refl.y + m0.z * u.y, refl.z + m0.z * u.z};
    return {m1.x, m1.y, m1.z};
  }
  if (kType == RT2_MAT_DIELECTRIC) {
    // Material.cpp:29-45
    float ri = front_face ? (1.0f / m0.w) : m0.w;
    F3 unit_dir = normalize3(d_in);
    float cos_theta = fminf(dot3(vneg(unit_dir), n), 1.0f);
    float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    bool cannot_refract = ri * sin_theta > 1.0f;
    float r0 = (1.0f - ri) / (1.0f + ri);
    r0 = r0 * r0;
    float om = 1.0f - cos_theta;
    float schlick = r0 + (1.0f - r0) * (om * om * om * om * om);
    if (cannot_refract || schlick > u01(r.z)) {
      dir = reflect3(unit_dir, n);
    } else {
      dir = refract3(unit_dir, n, ri);
    }
    return {1.0f, 1.0f, 1.0f};
  }
  // RT2_MAT_ISOTROPIC, Material.cpp:76-83
  dir = unit_vector(u01(r.x), u01(r.y));
  return texture_value(S, __float_as_uint(m0.y), p, tu, tv);
}

}  // namespace rt2dev
