// b2q_decode.cuh — definitions shared by the decode-tier kernels (b2q_decode.cu, b2q_decode2.cu): the per-warp
// cp.async.bulk ring geometry, the mma.sync wrapper, the per-unit scale / zero registers and the multi-set ("sibling"
// QuantLinears) tile index space.
#pragma once
#include "b2q_common.cuh"

namespace b2q {

constexpr int DEC_MAX_WARPS = 16;
constexpr int DEC_MAXM = 8;

template <typename T>
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma_16816<__half>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// One "quad" = 128 k x 32 features = two contiguous 1 KB pieces of T4 (k-blocks 2q and 2q+1, feature tiles 2nt and
// 2nt+1).  Each warp owns a private ring of DEC_STAGES quad buffers in shared memory, filled by cp.async.bulk
// (UBLKCP) and tracked by one mbarrier per stage: up to DEC_STAGES * 2 KB per warp are in flight with no register
// cost, which is what keeps > 100 KB per SM outstanding (the first, register-prefetch version of this kernel kept
// 2 KB per warp in flight and topped out at ~3 TB/s incremental: profiles/r01_decode_notes.md).
constexpr int DEC_STAGES = 4;  // maximum ring depth; the launch picks 4 or 2 stages (`stl` = log2) to fit shared memory
constexpr int DEC_QUAD_BYTES = 2048;

template <bool ASYM, bool G64>
struct DScale {
  uint16_t s[G64 ? 2 : 1][4];  // [group in quad][ftl * 2 + hi]
  uint32_t zw[(ASYM ? 1 : 0) * (G64 ? 2 : 1) + (ASYM ? 0 : 1)][4];
};

__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ float2 lds_f2(uint32_t a) {
  float2 r;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(a));
  return r;
}

__device__ __forceinline__ void issue_quad(uint32_t dst, uint32_t bar, const uint4* __restrict__ src,
                                           size_t kb_stride) {
  mbar_expect_tx(bar, DEC_QUAD_BYTES);
  bulk_load(dst, src, 1024, bar);                     // k-block 2q   : feature tiles 2nt, 2nt+1
  bulk_load(dst + 1024, src + kb_stride, 1024, bar);  // k-block 2q+1
}

// Persistent-style CTA: blockIdx.x strides over the 32-feature tiles (tile = blockIdx.x + i * gridDim.x), blockIdx.y is
// the split-K rank inside the cluster.  x is staged ONCE per CTA; the warps split the k-quads of every tile; each
// warp's bulk-copy ring runs ahead across tile boundaries.
// Up to DEC_MAX_SETS weight sets that consume the SAME activations (q/k/v, gate/up: "sibling" QuantLinears) are served
// by one launch: the 32-feature tiles of all sets form one index space (tile_end = running totals).
constexpr int DEC_MAX_SETS = 3;
struct DecSets {
  int nsets;
  int tile_end[DEC_MAX_SETS];
  int N[DEC_MAX_SETS];
  const uint4* packed[DEC_MAX_SETS];
  const void* scales[DEC_MAX_SETS];
  const uint32_t* qzeros[DEC_MAX_SETS];
  const void* bias[DEC_MAX_SETS];
  void* out[DEC_MAX_SETS];
};

template <typename T>
struct TileRef {
  const uint4* w;
  const T* sc;
  const uint32_t* zq;
  const T* bias;
  T* out;
  int N, nt;
};

template <typename T>
__device__ __forceinline__ TileRef<T> resolve_tile(const DecSets& S, int gt) {
  int s = 0, start = 0;
  if (S.nsets > 1 && gt >= S.tile_end[0]) {
    s = 1;
    start = S.tile_end[0];
    if (S.nsets > 2 && gt >= S.tile_end[1]) {
      s = 2;
      start = S.tile_end[1];
    }
  }
  TileRef<T> r;
  r.w = S.packed[s];
  r.sc = reinterpret_cast<const T*>(S.scales[s]);
  r.zq = S.qzeros[s];
  r.bias = reinterpret_cast<const T*>(S.bias[s]);
  r.out = reinterpret_cast<T*>(S.out[s]);
  r.N = S.N[s];
  r.nt = gt - start;
  return r;
}

}  // namespace b2q
