// b2q_dequant.cuh — exact (q - z) * s dequantisation of packed B2Q words for the tensor-core tiers.
#pragma once
#include "b2q_common.cuh"

namespace b2q {

struct SZRaw {
  uint32_t s;   // scale, 16-bit payload
  uint32_t zw;  // packed zero word (or unused)
};

template <typename T, int BITS, bool ASYM>
__device__ __forceinline__ SZRaw load_sz(const T* __restrict__ scales, const uint32_t* __restrict__ qzeros, int g,
                                         int n, int N) {
  SZRaw r;
  r.s = *reinterpret_cast<const uint16_t*>(scales + (size_t)g * N + n);
  r.zw = 0;
  if (ASYM) {
    constexpr int PF = 32 / BITS;
    r.zw = qzeros[(size_t)g * (N / PF) + n / PF];
  }
  return r;
}

// exact dequant of one packed uint4 (32 k of one feature for 4-bit, 16 k for 8-bit) into K-consecutive
// 16-byte groups; out[i] holds 8 consecutive k.
template <typename T, int BITS>
struct Dequant;

// 4-bit fragment-major uint4 (see b2q_common.cuh): features (g, g+8) x 16 consecutive k.
// lo[c] / hi[c] = 8 consecutive k (chunk c = 0,1 of the lane's 16) of feature g / g+8, exactly (q - z) * s.
template <>
struct Dequant<__half, 4> {
  __device__ static __forceinline__ void run(const uint4& pv, uint32_t slo16, int zlo_i, uint32_t shi16, int zhi_i,
                                             uint4 (&lo)[2], uint4 (&hi)[2]) {
    const uint32_t slu = slo16 | (slo16 << 16), shu = shi16 | (shi16 << 16);
    const __half2 sl = *reinterpret_cast<const __half2*>(&slu), sh = *reinterpret_cast<const __half2*>(&shu);
    const __half2 zlo = __float2half2_rn(1024.f + (float)zlo_i);  // exact
    const __half2 zhi = __float2half2_rn(-(64.f + (float)zhi_i));  // exact
    const __half2 sixteenth = __float2half2_rn(0.0625f);
    const uint32_t w[4] = {pv.x, pv.y, pv.z, pv.w};
    uint32_t l[8], u[8];
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
      uint32_t h[4];
      ET<__half>::unpack_w4(w[s4], h);
      __half2 v0 = __hmul2(__hsub2(*reinterpret_cast<__half2*>(&h[0]), zlo), sl);
      __half2 v1 = __hmul2(__hfma2(*reinterpret_cast<__half2*>(&h[1]), sixteenth, zhi), sh);
      __half2 v2 = __hmul2(__hsub2(*reinterpret_cast<__half2*>(&h[2]), zlo), sl);
      __half2 v3 = __hmul2(__hfma2(*reinterpret_cast<__half2*>(&h[3]), sixteenth, zhi), sh);
      l[2 * s4] = *reinterpret_cast<uint32_t*>(&v0);
      l[2 * s4 + 1] = *reinterpret_cast<uint32_t*>(&v2);
      u[2 * s4] = *reinterpret_cast<uint32_t*>(&v1);
      u[2 * s4 + 1] = *reinterpret_cast<uint32_t*>(&v3);
    }
    lo[0] = make_uint4(l[0], l[1], l[2], l[3]);
    lo[1] = make_uint4(l[4], l[5], l[6], l[7]);
    hi[0] = make_uint4(u[0], u[1], u[2], u[3]);
    hi[1] = make_uint4(u[4], u[5], u[6], u[7]);
  }
};

template <>
struct Dequant<__nv_bfloat16, 4> {
  __device__ static __forceinline__ void run(const uint4& pv, uint32_t slo16, int zlo_i, uint32_t shi16, int zhi_i,
                                             uint4 (&lo)[2], uint4 (&hi)[2]) {
    const uint32_t slu = slo16 | (slo16 << 16), shu = shi16 | (shi16 << 16);
    const __nv_bfloat162 sl = *reinterpret_cast<const __nv_bfloat162*>(&slu);
    const __nv_bfloat162 sh = *reinterpret_cast<const __nv_bfloat162*>(&shu);
    const __nv_bfloat162 zl = __float2bfloat162_rn(128.f + (float)zlo_i);  // exact (<= 143)
    const __nv_bfloat162 zh = __float2bfloat162_rn(128.f + (float)zhi_i);
    const uint32_t w[4] = {pv.x, pv.y, pv.z, pv.w};
    uint32_t l[8], u[8];
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
      uint32_t h[4];
      ET<__nv_bfloat16>::unpack_w4(w[s4], h);
      __nv_bfloat162 v0 = __hmul2(__hsub2(*reinterpret_cast<__nv_bfloat162*>(&h[0]), zl), sl);
      __nv_bfloat162 v1 = __hmul2(__hsub2(*reinterpret_cast<__nv_bfloat162*>(&h[1]), zh), sh);
      __nv_bfloat162 v2 = __hmul2(__hsub2(*reinterpret_cast<__nv_bfloat162*>(&h[2]), zl), sl);
      __nv_bfloat162 v3 = __hmul2(__hsub2(*reinterpret_cast<__nv_bfloat162*>(&h[3]), zh), sh);
      l[2 * s4] = *reinterpret_cast<uint32_t*>(&v0);
      l[2 * s4 + 1] = *reinterpret_cast<uint32_t*>(&v2);
      u[2 * s4] = *reinterpret_cast<uint32_t*>(&v1);
      u[2 * s4 + 1] = *reinterpret_cast<uint32_t*>(&v3);
    }
    lo[0] = make_uint4(l[0], l[1], l[2], l[3]);
    lo[1] = make_uint4(l[4], l[5], l[6], l[7]);
    hi[0] = make_uint4(u[0], u[1], u[2], u[3]);
    hi[1] = make_uint4(u[4], u[5], u[6], u[7]);
  }
};

template <>
struct Dequant<__half, 8> {
  // 16 k per uint4 -> 2 x uint4
  __device__ static __forceinline__ void run(const uint4& pv, uint32_t s16, int z, uint4 (&o)[2]) {
    const uint32_t s2u = s16 | (s16 << 16);
    const __half2 s2 = *reinterpret_cast<const __half2*>(&s2u);
    const __half2 zb = __float2half2_rn(1024.f + (float)z);  // exact (<= 1279)
    const uint32_t w[4] = {pv.x, pv.y, pv.z, pv.w};
    uint32_t r[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      uint32_t p0 = __byte_perm(w[t], 0x64006400u, 0x7150);
      uint32_t p1 = __byte_perm(w[t], 0x64006400u, 0x7352);
      __half2 v0 = __hmul2(__hsub2(*reinterpret_cast<__half2*>(&p0), zb), s2);
      __half2 v1 = __hmul2(__hsub2(*reinterpret_cast<__half2*>(&p1), zb), s2);
      r[2 * t] = *reinterpret_cast<uint32_t*>(&v0);
      r[2 * t + 1] = *reinterpret_cast<uint32_t*>(&v1);
    }
    o[0] = make_uint4(r[0], r[1], r[2], r[3]);
    o[1] = make_uint4(r[4], r[5], r[6], r[7]);
  }
};

template <>
struct Dequant<__nv_bfloat16, 8> {
  __device__ static __forceinline__ void run(const uint4& pv, uint32_t s16, int z, uint4 (&o)[2]) {
    const float s = __uint_as_float(s16 << 16);
    const uint32_t w[4] = {pv.x, pv.y, pv.z, pv.w};
    uint32_t r[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      // (q - z) exact in fp32, product with the bf16 scale exact in fp32, ONE rounding to bf16
      const float q0 = (float)((int)(w[t] & 0xFFu) - z), q1 = (float)((int)((w[t] >> 8) & 0xFFu) - z);
      const float q2 = (float)((int)((w[t] >> 16) & 0xFFu) - z), q3 = (float)((int)(w[t] >> 24) - z);
      r[2 * t] = ET<__nv_bfloat16>::pack2(q0 * s, q1 * s);
      r[2 * t + 1] = ET<__nv_bfloat16>::pack2(q2 * s, q3 * s);
    }
    o[0] = make_uint4(r[0], r[1], r[2], r[3]);
    o[1] = make_uint4(r[4], r[5], r[6], r[7]);
  }
};


}  // namespace b2q
