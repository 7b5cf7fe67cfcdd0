// b2q_gemm2.cuh — constants and PTX wrappers shared by the CTA-pair (cta_group::2) prefill kernels
// (b2q_gemm2.cu: one tile per pair / persistent pairs).
#pragma once
#include <cuda.h>

#include "b2q_common.cuh"

namespace b2q {

constexpr int G2_BK = 64;
constexpr int G2_THREADS_MAX = 64 + 8 * 32;
constexpr int G2_STAGES = 5;
constexpr int G2_A_BYTES = 128 * G2_BK * 2;   // 16 KB: this CTA's 128 token rows
constexpr int G2_B_BYTES = 128 * G2_BK * 2;   // 16 KB: this CTA's 128 feature rows (dequantised)
constexpr int G2_P_BYTES = 4096 + 1024;       // packed int4 block of the same 128 features x 64 k + scale/zero rows
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES + G2_P_BYTES;
constexpr int G2_SMEM_BYTES = G2_STAGES * G2_STAGE_BYTES + 512 + 1024;
constexpr int G2_TMEM_COLS = 256;

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics on purpose: the .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR per arrive and was 43 %
  // of all stall samples (profiles/r01_gemm2_r1a.txt); ordering of the dequantised tile towards the tensor cores is
  // given by fence.proxy.async + the CTA-wide named barrier that precedes this single arrive
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t dst, const void* tmap, uint32_t bar_cluster, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(dst),
      "l"(tmap), "r"(bar_cluster), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_f16_cg2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{ .reg .pred p; setp.ne.b32 p, %4, 0;\n"
      "  tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_cg2_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}

}  // namespace b2q
