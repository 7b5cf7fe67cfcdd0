// b2q_gemm.cuh — tile geometry shared by the single-CTA tcgen05 kernels (b2q_gemm.cu).
#pragma once
#include "b2q_common.cuh"

namespace b2q {

constexpr int G_BN = 128;
constexpr int G_BK = 64;
constexpr int G_THREADS = 192;
constexpr int G_DQ_THREADS = 128;

template <int BITS, int MT, int STAGES>
struct GemmCfg {
  static constexpr int SUB = BITS / 4;
  static constexpr int A_BYTES = MT * 128 * G_BK * 2;
  static constexpr int B_BYTES = G_BN * G_BK * 2;
  static constexpr int P_CHUNK_BYTES = 4 * SUB * 512;  // 8-bit: 4 feature tiles (of 32) x 32 k
  // 128 features x 64 k of packed codes (4-bit: one 4 KB T4 row block) + for 4-bit the scale / zero-point rows of
  // the (up to 2) groups of the block: 2 x (256 B scales + 64 B packed zeros), staged by the producer
  static constexpr int P_BYTES = 2 * P_CHUNK_BYTES + (BITS == 4 ? 1024 : 0);
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES + P_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;
  static constexpr int TMEM_COLS = MT * G_BN;
};

}  // namespace b2q
