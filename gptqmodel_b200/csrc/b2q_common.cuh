// b2q_common.cuh — shared device helpers for the B200 (sm_100a) GPTQ QuantLinear kernels.
//
// Prepacked weight layouts ("B2Q tiles", produced by b2q_prepack.cu from the checkpoint layout
// qweight int32 [K*bits/32, N] of /root/reference/gptqmodel/nn_modules/qlinear/__init__.py:827-865):
//
//   4-bit:  uint4 T4[K/64][N/16][32]      one warp-wide 512-byte row = 16 output features x 64 k
//   8-bit:  uint4 T8[K/32][N/32][2][32]   one uint4 = 16 consecutive k of ONE output feature n
//
// 4-bit ("fragment-major"): lane = 4*g + t (g = 0..7, t = 0..3) owns features 16*ft+g and 16*ft+g+8 and the 16
// consecutive k  64*kb + 16*t .. +15.  Word s (0..3) of its uint4 covers k = 64*kb + 16*t + 4*s + {0,1,2,3};
// inside a word the nibble at bits [4p, 4p+4) holds
//      p = 0,4 : feature g   , k+0, k+1        p = 1,5 : feature g+8 , k+0, k+1
//      p = 2,6 : feature g   , k+2, k+3        p = 3,7 : feature g+8 , k+2, k+3
// so that  (w & 0x000f000f)|0x64006400 = half2(1024+q) of (g, k+0..1),   (w & 0x00f000f0)|0x64006400 =
// half2(1024+16q) of (g+8, k+0..1), and the same two masks on (w >> 8) give k+2..3: FOUR LOP3 (+1 shift) turn a
// word into the four A-operand registers of one mma.sync.m16n8k16 (decode tier: rows g / g+8, the k permutation is
// shared with the activation fragment), and the same registers are K-consecutive pairs of ONE feature row, so the
// tcgen05 GEMM tier stores them as 16-byte K-major shared-memory chunks.  Rows g carry the magic bias 1024, rows g+8
// carry 16*(64+q): one bias class per output row, removed once per group with the activation sum.
// A CTA tile of 128 features x 64 k is ONE contiguous 4 KB block for cp.async.bulk.
// With act-order the rows are first sorted by group (k' -> original row perm[k']) so groups are contiguous.
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2q {

// ------------------------------------------------------------------------------------------------
// small PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ uint32_t lop3_and_or(uint32_t a, uint32_t mask, uint32_t orv) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(mask), "r"(orv));  // (a & b) | c
  return r;
}

// d = a(f16 half of reg, selected by H) * b(f16 half) + c(f32)  -> SASS FHFMA on sm_100a
__device__ __forceinline__ float fhfma_lo(uint32_t a2, uint32_t b2, float c) {
  float d;
  asm("{ .reg .b16 al, ah, bl, bh; mov.b32 {al, ah}, %1; mov.b32 {bl, bh}, %2;\n"
      "  fma.rn.f32.f16 %0, al, bl, %3; }"
      : "=f"(d)
      : "r"(a2), "r"(b2), "f"(c));
  return d;
}
__device__ __forceinline__ float fhfma_hi(uint32_t a2, uint32_t b2, float c) {
  float d;
  asm("{ .reg .b16 al, ah, bl, bh; mov.b32 {al, ah}, %1; mov.b32 {bl, bh}, %2;\n"
      "  fma.rn.f32.f16 %0, ah, bh, %3; }"
      : "=f"(d)
      : "r"(a2), "r"(b2), "f"(c));
  return d;
}
__device__ __forceinline__ float fbfma_lo(uint32_t a2, uint32_t b2, float c) {
  float d;
  asm("{ .reg .b16 al, ah, bl, bh; mov.b32 {al, ah}, %1; mov.b32 {bl, bh}, %2;\n"
      "  fma.rn.f32.bf16 %0, al, bl, %3; }"
      : "=f"(d)
      : "r"(a2), "r"(b2), "f"(c));
  return d;
}
__device__ __forceinline__ float fbfma_hi(uint32_t a2, uint32_t b2, float c) {
  float d;
  asm("{ .reg .b16 al, ah, bl, bh; mov.b32 {al, ah}, %1; mov.b32 {bl, bh}, %2;\n"
      "  fma.rn.f32.bf16 %0, ah, bh, %3; }"
      : "=f"(d)
      : "r"(a2), "r"(b2), "f"(c));
  return d;
}

// ---- mbarrier ----------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- TMA / bulk copies -------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// ---- tcgen05 -----------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::f16 (fp16/bf16 in, fp32 accumulate)
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{ .reg .pred p; setp.ne.b32 p, %4, 0;\n"
      "  tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
        "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
        "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B, rows of 128 bytes, 8-row atoms 1024 B apart.
// Bit layout (cute/arch/mma_sm100_desc.hpp SmemDescriptor): [0,14) addr>>4, [16,30) LBO>>4,
// [32,46) SBO>>4, [46,48) version=1, [61,64) layout (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: c=f32, a/b format (0 f16, 1 bf16), both K-major, N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int fmt, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}

// ---- bounded spin on a flag another GPU writes (peer-memory all-reduce kernels) -------------------
// A dead or wedged peer must not hang this GPU for ever: the wait gives up after
// B2Q_PEER_TIMEOUT_NS of %globaltimer and the caller records 1 + peer rank in its status word.
constexpr unsigned long long B2Q_PEER_TIMEOUT_NS = 2000000000ull;  // 2 s
__device__ __forceinline__ bool spin_until_geq_sys(const uint32_t* p, uint32_t target) {
  unsigned long long t0 = 0;
  for (unsigned it = 0;; ++it) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if ((int32_t)(v - target) >= 0) return true;
    if ((it & 1023u) == 1023u) {
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > B2Q_PEER_TIMEOUT_NS) return false;
    }
  }
}

// ---- cluster helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_saddr, uint32_t cta) {
  uint32_t ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_saddr), "r"(cta));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}

// ------------------------------------------------------------------------------------------------
// element-type traits (fp16 / bf16 activations)
// ------------------------------------------------------------------------------------------------
template <typename T>
struct ET;
template <>
struct ET<__half> {
  static constexpr int FMT = 0;
  static constexpr uint32_t MAGIC = 0x64006400u;  // half2(1024, 1024)
  static constexpr float LO_BASE = 1024.f;        // nibble in mantissa bits [0,4)
  static constexpr float HI_BASE = 64.f;          // nibble in mantissa bits [4,8): (1024+16q)/16 = 64+q
  static constexpr float HI_SCALE = 0.0625f;
  // one 32-bit word -> 4 packed pairs: h[0] = (g, k0..1), h[1] = (g+8, k0..1), h[2] = (g, k2..3), h[3] = (g+8, k2..3)
  // h[0], h[2] = LO_BASE + q ; h[1], h[3] = (HI_BASE + q) / HI_SCALE      (all exact)
  __device__ static __forceinline__ void unpack_w4(uint32_t w, uint32_t (&h)[4]) {
    h[0] = lop3_and_or(w, 0x000f000fu, MAGIC);
    h[1] = lop3_and_or(w, 0x00f000f0u, MAGIC);
    const uint32_t w8 = w >> 8;
    h[2] = lop3_and_or(w8, 0x000f000fu, MAGIC);
    h[3] = lop3_and_or(w8, 0x00f000f0u, MAGIC);
  }
  __device__ static __forceinline__ float to_f(__half v) { return __half2float(v); }
  __device__ static __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
  __device__ static __forceinline__ float fma_lo(uint32_t a, uint32_t b, float c) { return fhfma_lo(a, b, c); }
  __device__ static __forceinline__ float fma_hi(uint32_t a, uint32_t b, float c) { return fhfma_hi(a, b, c); }
  __device__ static __forceinline__ uint32_t pack2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};
template <>
struct ET<__nv_bfloat16> {
  static constexpr int FMT = 1;
  static constexpr uint32_t MAGIC = 0x43004300u;  // bf16x2(128, 128); 7 mantissa bits
  static constexpr float LO_BASE = 128.f;         // nibble in mantissa bits [0,4)
  static constexpr float HI_BASE = 128.f;         // bf16 has no room for a nibble at bits [4,8): shift instead
  static constexpr float HI_SCALE = 1.0f;
  __device__ static __forceinline__ void unpack_w4(uint32_t w, uint32_t (&h)[4]) {
    h[0] = lop3_and_or(w, 0x000f000fu, MAGIC);
    h[1] = lop3_and_or(w >> 4, 0x000f000fu, MAGIC);
    h[2] = lop3_and_or(w >> 8, 0x000f000fu, MAGIC);
    h[3] = lop3_and_or(w >> 12, 0x000f000fu, MAGIC);
  }
  __device__ static __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __forceinline__ __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
  __device__ static __forceinline__ float fma_lo(uint32_t a, uint32_t b, float c) { return fbfma_lo(a, b, c); }
  __device__ static __forceinline__ float fma_hi(uint32_t a, uint32_t b, float c) { return fbfma_hi(a, b, c); }
  __device__ static __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
};

}  // namespace b2q
