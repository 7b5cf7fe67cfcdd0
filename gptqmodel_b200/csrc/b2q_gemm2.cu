// b2q_gemm2.cu — prefill tier on CTA PAIRS: out[M, N] = x[M, K] @ dequant(W4)[K, N] with tcgen05.mma.cta_group::2.
//
// Why pairs: the 1-CTA tier (b2q_gemm.cu) is bound by SHARED-MEMORY BANDWIDTH, not by the tensor pipe — per 64-k block
// a 256x128 CTA tile moves 64 KB of UMMA operand reads + 16 KB of dequant stores + 36 KB of TMA writes through the
// 128 B/clk shared memory in the 512 clk the MMAs need (234 B/clk wanted -> ~55 % of peak, which is what was
// measured: profiles/r01_gemm_notes.md).  With cta_group::2 a pair of SMs computes a 256 (tokens) x 256 (features)
// tile; each CTA stages only ITS 128 token rows of A and dequantises only ITS 128 feature rows of B, the tensor cores
// of both SMs read both halves: 32 KB operand reads + 16 KB dequant stores + 20 KB TMA per 512 clk per SM (~144 B/clk).
//
//   cluster (2,1,1); rank r = %cluster_ctarank owns tokens m0+128r.. and features n0+128r..
//   warp 0      producer : TMA (cp.async.bulk.tensor .cta_group::2) of its A half, signalling the LEADER's mbarrier;
//                          cp.async.bulk of its 4 KB packed-weight block on a local mbarrier
//   warp 1      MMA      : both CTAs allocate TMEM (cta_group::2); the leader's elected lane issues
//                          tcgen05.mma.cta_group::2.kind::f16 (M=256, N=256, K=16), tcgen05.commit multicast to both
//   warps 2..5  dequant  : as in the 1-CTA tier; completion arrives on the leader's mbarrier (remote arrive);
//                          afterwards epilogue of the CTA's own 128 x 256 accumulator (TMEM -> regs -> global)
#include <cuda.h>

#include <cstdlib>

#include "b2q_common.cuh"
#include "b2q_dequant.cuh"
#include "b2q_gemm2.cuh"
#include "b2q_internal.h"

namespace b2q {

// DQW = number of dequant (+ epilogue) warps: 4 (two packed uint4 per thread and stage) or 8 (one)
template <typename T, bool ASYM, int DQW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + DQW * 32, 1)
    gemm2_kernel(const __grid_constant__ CUtensorMap tmap_x, const uint4* __restrict__ packed,
                 const T* __restrict__ scales, const uint32_t* __restrict__ qzeros, const T* __restrict__ bias,
                 T* __restrict__ out, int M, int K, int N, int group_size, int gshc) {
  using E = ET<T>;
  constexpr int STAGES = G2_STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t sA = smem_base;
  const uint32_t sB = sA + STAGES * G2_A_BYTES;
  const uint32_t sP = sB + STAGES * G2_B_BYTES;
  const uint32_t sBar = sP + STAGES * G2_P_BYTES;
  // leader-owned (used through cluster addresses by the peer): fullA, bready.  per-CTA: fullP, empty, tfull
  const uint32_t bar_fullA = sBar, bar_bready = sBar + 8 * STAGES, bar_fullP = sBar + 16 * STAGES;
  const uint32_t bar_empty = sBar + 24 * STAGES, bar_tfull = sBar + 32 * STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + (sBar - smem_base) + 32 * STAGES + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int FT = N >> 4;
  const int n0 = (blockIdx.x >> 1) * 256 + (int)rank * 128;  // this CTA's 128 feature rows
  const int npair0 = (blockIdx.x >> 1) * 256;               // the pair's 256 output columns
  const int m0 = blockIdx.y * 256 + (int)rank * 128;         // this CTA's 128 token rows
  const int ft0 = n0 >> 4;
  const int nkb = K / G2_BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_fullA + 8 * s, 1);     // leader: its own arrive.expect_tx (both CTAs' TMA bytes)
      mbar_init(bar_bready + 8 * s, 2);    // leader: ONE aggregated arrive per CTA (remote arrives are slow)
      mbar_init(bar_fullP + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(G2_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;

  if (warp == 0) {
    // ================================ producer ================================
    if (lane == 0) {
      const int nft = max(0, min(8, FT - ft0));
      const uint32_t pbytes = (uint32_t)nft * 512u;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        if (rank == 0) mbar_expect_tx(bar_fullA + 8 * s, 2 * G2_A_BYTES);
        tma_load_2d_cg2(sA + s * G2_A_BYTES, &tmap_x, mapa_u32(bar_fullA + 8 * s, 0), kb * G2_BK, m0);
        if (pbytes > 0) {
          const int g0 = (2 * kb) >> gshc, g1 = (2 * kb + 1) >> gshc;
          const int nrows = (g1 != g0) ? 2 : 1;
          const uint32_t sbytes = (uint32_t)min(128, N - n0) * 2u, zbytes = ASYM ? sbytes / 4u : 0u;
          mbar_expect_tx(bar_fullP + 8 * s, pbytes + nrows * (sbytes + zbytes));
          bulk_load(sP + s * G2_P_BYTES, packed + ((size_t)kb * FT + ft0) * 32, pbytes, bar_fullP + 8 * s);
          for (int r = 0; r < nrows; ++r) {
            const int gr = r ? g1 : g0;
            bulk_load(sP + s * G2_P_BYTES + 4096 + r * 320, scales + (size_t)gr * N + n0, sbytes, bar_fullP + 8 * s);
            if (ASYM)
              bulk_load(sP + s * G2_P_BYTES + 4096 + r * 320 + 256, qzeros + (size_t)gr * (N >> 3) + (n0 >> 3),
                        zbytes, bar_fullP + 8 * s);
          }
        } else {
          mbar_arrive(bar_fullP + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ================================
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(E::FMT, 256, 256);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(bar_fullA + 8 * s, ph);
        mbar_wait(bar_bready + 8 * s, ph);
        tc_fence_after();
        if (lane == 0) {
          const uint64_t adesc = umma_desc_k_sw128(sA + s * G2_A_BYTES);
          const uint64_t bdesc = umma_desc_k_sw128(sB + s * G2_B_BYTES);
#pragma unroll
          for (int k = 0; k < G2_BK / 16; ++k)
            umma_f16_cg2(tbase, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          umma_commit_cg2_mc(bar_empty + 8 * s, 3);
          if (kb == nkb - 1) umma_commit_cg2_mc(bar_tfull, 3);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================ dequant warps ================================
    const int t = threadIdx.x - 64;  // 0 .. DQW*32-1
    constexpr int NU = 8 / DQW;      // packed uint4 per thread and stage
    constexpr int ZSYM = 8;
    const int lp = t & 31, g = lp >> 2, tt = lp & 3;
    int f[2 * NU];  // tile-local feature rows (lo / hi) of this thread's uint4(s)
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      f[2 * u] = ((t + u * (DQW * 32)) >> 5) * 16 + g;
      f[2 * u + 1] = f[2 * u] + 8;
    }
    const uint32_t bready_leader = mapa_u32(bar_bready, 0);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(bar_fullP + 8 * s, ph);
      const uint8_t* pst = smem + (sP - smem_base) + s * G2_P_BYTES;
      const uint4* pj = reinterpret_cast<const uint4*>(pst);
      const int grow = ((2 * kb + (tt >> 1)) >> gshc) - ((2 * kb) >> gshc);  // 0 or 1
      const uint8_t* srow = pst + 4096 + grow * 320;
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const uint4 pv = pj[t + u * (DQW * 32)];
        const uint32_t s_lo = *reinterpret_cast<const uint16_t*>(srow + f[2 * u] * 2);
        const uint32_t s_hi = *reinterpret_cast<const uint16_t*>(srow + f[2 * u + 1] * 2);
        int zl = ZSYM, zh = ZSYM;
        if (ASYM) {
          const uint32_t zwl = *reinterpret_cast<const uint32_t*>(srow + 256 + (f[2 * u] >> 3) * 4);
          const uint32_t zwh = *reinterpret_cast<const uint32_t*>(srow + 256 + (f[2 * u + 1] >> 3) * 4);
          zl = (int)((zwl >> (4 * g)) & 15u);
          zh = (int)((zwh >> (4 * g)) & 15u);
        }
        uint4 lo[2], hi[2];
        Dequant<T, 4>::run(pv, s_lo, zl, s_hi, zh, lo, hi);
        const uint32_t sw = (uint32_t)g;
        const uint32_t rlo = sB + s * G2_B_BYTES + f[2 * u] * 128;
        const uint32_t rhi = rlo + 8 * 128;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t off = (((uint32_t)(2 * tt + c)) ^ sw) << 4;
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(rlo + off), "r"(lo[c].x), "r"(lo[c].y),
                       "r"(lo[c].z), "r"(lo[c].w)
                       : "memory");
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(rhi + off), "r"(hi[c].x), "r"(hi[c].y),
                       "r"(hi[c].z), "r"(hi[c].w)
                       : "memory");
        }
      }
      fence_proxy_async_smem();
      // aggregate the 128 dequant threads on a named barrier, then ONE (possibly remote) arrive per CTA: 128
      // individual remote mbarrier arrives per stage serialised on the cluster network and cost more than the MMAs
      asm volatile("bar.sync 1, %0;" ::"n"(DQW * 32) : "memory");
      if (t == 0) mbar_arrive_cluster(bready_leader + 8 * s);
    }

    // ================================ epilogue: own 128 tokens x 256 features ================================
    mbar_wait(bar_tfull, 0);
    tc_fence_after();
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    constexpr int CPW = 8 / (DQW / 4);             // 32-column chunks per warp
    const int cc0 = ((warp - 2) >> 2) * CPW;       // with 8 warps two warps share a TMEM lane quarter
#pragma unroll 1
    for (int cc = cc0; cc < cc0 + CPW; ++cc) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tbase + ((uint32_t)(q * 32) << 16) + cc * 32, r);
      tmem_ld_wait();
      const int nc = npair0 + cc * 32;
      if (row < M && nc < N) {
        T* dst = out + (size_t)row * N + nc;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          uint32_t pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float f0 = __uint_as_float(r[v * 8 + 2 * i]), f1 = __uint_as_float(r[v * 8 + 2 * i + 1]);
            if (bias != nullptr) {
              f0 = E::to_f(E::from_f(f0)) + E::to_f(bias[nc + v * 8 + 2 * i]);
              f1 = E::to_f(E::from_f(f1)) + E::to_f(bias[nc + v * 8 + 2 * i + 1]);
            }
            pk[i] = E::pack2(f0, f1);
          }
          *reinterpret_cast<uint4*>(dst + v * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync_all();  // both CTAs done with TMEM / each other's shared memory
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(G2_TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent variant: one CTA pair per SM pair loops over the 256x256 output tiles (static round-robin, M fastest so
// the pairs that run concurrently share a weight tile in L2); TMEM holds TWO 256-column accumulators so that a
// dedicated group of 4 epilogue warps drains tile i while the mainloop already runs tile i+1; barrier init, TMEM
// allocation, cluster syncs and the pipeline fill are paid once per CTA instead of once per tile (the per-tile fixed
// cost of the non-persistent kernel was measured at ~7 us, profiles/r01_gemm_notes.md).
//   warp 0 activation producer | warp 1 MMA (leader) | warps 2..5 dequant | warps 6..9 epilogue | warp 10 weight producer
// (two producer threads: a single thread issuing the activation TMA AND the packed / scale / zero bulk copies of every
//  k-block needs ~150 clk per asynchronous copy, 4-5 per block = the 770 clk per block this kernel was measured at against
//  a 512 clk MMA floor — the same wall the small-batch tier hit, profiles/r02_midm_notes.md)
constexpr int G2P_THREADS = 352;

// Sibling QuantLinears that consume the SAME activations (q|k|v, gate|up) share ONE persistent launch: the 256-feature tile
// COLUMNS of all sets form one index space (tn_end = running totals), so the 74 CTA pairs see 192 / 896 tiles instead of
// 128 + 32 + 32 / 448 + 448 — fewer partially filled waves and two launches less per layer.
constexpr int G2_MAX_SETS = 3;
struct G2Sets {
  int nsets;
  int tn_end[G2_MAX_SETS];
  int N[G2_MAX_SETS];
  const uint4* packed[G2_MAX_SETS];
  const void* scales[G2_MAX_SETS];
  const uint32_t* qzeros[G2_MAX_SETS];
  const void* bias[G2_MAX_SETS];
  void* out[G2_MAX_SETS];
};
struct G2Tile {
  int set, tn;  // weight set and tile column inside it
};
__device__ __forceinline__ G2Tile g2_resolve(const G2Sets& S, int tn) {
  G2Tile t;
  t.set = 0;
  t.tn = tn;
  if (S.nsets > 1 && tn >= S.tn_end[0]) {
    t.set = 1;
    t.tn = tn - S.tn_end[0];
    if (S.nsets > 2 && tn >= S.tn_end[1]) {
      t.set = 2;
      t.tn = tn - S.tn_end[1];
    }
  }
  return t;
}
constexpr int G2P_TMEM_COLS = 512;

template <typename T, bool ASYM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2P_THREADS, 1)
    gemm2p_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ G2Sets S, int M, int K,
                  int group_size, int gshc, int TM, int ntiles_total) {
  using E = ET<T>;
  constexpr int STAGES = G2_STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t sA = smem_base;
  const uint32_t sB = sA + STAGES * G2_A_BYTES;
  const uint32_t sP = sB + STAGES * G2_B_BYTES;
  const uint32_t sBar = sP + STAGES * G2_P_BYTES;
  const uint32_t bar_fullA = sBar, bar_bready = sBar + 8 * STAGES, bar_fullP = sBar + 16 * STAGES;
  const uint32_t bar_empty = sBar + 24 * STAGES;
  const uint32_t bar_tfull = sBar + 32 * STAGES;        // [2] per CTA (multicast commit)
  const uint32_t bar_tempty = sBar + 32 * STAGES + 16;  // [2] leader-owned: both CTAs' epilogues drained buffer a
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + (sBar - smem_base) + 32 * STAGES + 40);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int nkb = K / G2_BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    for (int s = 0; s < STAGES; ++s) {
      // ONE "operands ready" barrier per stage (leader-owned): the activation producer's arrive.expect_tx (both CTAs' TMA
      // halves complete_tx on it) + one arrive per CTA when its 128 feature rows are dequantised.  The MMA issuer waits
      // once per k-block instead of twice (its issue loop, not the tensor pipe, sets the k-block time: r02_midm_notes.md).
      mbar_init(bar_fullA + 8 * s, 3);
      mbar_init(bar_bready + 8 * s, 1);  // (unused)
      mbar_init(bar_fullP + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 2);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)),
                 "r"(G2P_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;

  if (warp == 0) {
    // ================================ activation producer ================================
    if (lane == 0) {
      int kbc = 0;  // k-block counter across tiles (stage / phase bookkeeping)
      for (int tile = pair; tile < ntiles_total; tile += npairs) {
        const int tm = tile % TM;
        const int m0 = tm * 256 + (int)rank * 128;
        for (int kb = 0; kb < nkb; ++kb, ++kbc) {
          const int s = kbc % STAGES;
          const uint32_t ph = (kbc / STAGES) & 1;
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          if (rank == 0) mbar_expect_tx(bar_fullA + 8 * s, 2 * G2_A_BYTES);
          tma_load_2d_cg2(sA + s * G2_A_BYTES, &tmap_x, mapa_u32(bar_fullA + 8 * s, 0), kb * G2_BK, m0);
        }
      }
    }
  } else if (warp == 10) {
    // ================================ weight producer ================================
    if (lane == 0) {
      int kbc = 0;
      for (int tile = pair; tile < ntiles_total; tile += npairs) {
        const G2Tile gt = g2_resolve(S, tile / TM);
        const int N = S.N[gt.set], FT = N >> 4;
        const uint4* packed = S.packed[gt.set];
        const T* scales = reinterpret_cast<const T*>(S.scales[gt.set]);
        const uint32_t* qzeros = S.qzeros[gt.set];
        const int n0 = gt.tn * 256 + (int)rank * 128;
        const int ft0 = n0 >> 4;
        const int nft = max(0, min(8, FT - ft0));
        const uint32_t pbytes = (uint32_t)nft * 512u;
        for (int kb = 0; kb < nkb; ++kb, ++kbc) {
          const int s = kbc % STAGES;
          const uint32_t ph = (kbc / STAGES) & 1;
          mbar_wait(bar_empty + 8 * s, ph ^ 1);
          if (pbytes > 0) {
            const int g0 = (2 * kb) >> gshc, g1 = (2 * kb + 1) >> gshc;
            const int nrows = (g1 != g0) ? 2 : 1;
            const uint32_t sbytes = (uint32_t)min(128, N - n0) * 2u, zbytes = ASYM ? sbytes / 4u : 0u;
            mbar_expect_tx(bar_fullP + 8 * s, pbytes + nrows * (sbytes + zbytes));
            bulk_load(sP + s * G2_P_BYTES, packed + ((size_t)kb * FT + ft0) * 32, pbytes, bar_fullP + 8 * s);
            for (int r = 0; r < nrows; ++r) {
              const int gr = r ? g1 : g0;
              bulk_load(sP + s * G2_P_BYTES + 4096 + r * 320, scales + (size_t)gr * N + n0, sbytes,
                        bar_fullP + 8 * s);
              if (ASYM)
                bulk_load(sP + s * G2_P_BYTES + 4096 + r * 320 + 256, qzeros + (size_t)gr * (N >> 3) + (n0 >> 3),
                          zbytes, bar_fullP + 8 * s);
            }
          } else {
            mbar_arrive(bar_fullP + 8 * s);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ================================
    if (rank == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(E::FMT, 256, 256);
      int kbc = 0, it = 0;
      for (int tile = pair; tile < ntiles_total; tile += npairs, ++it) {
        const int a = it & 1;
        mbar_wait(bar_tempty + 8 * a, (uint32_t)((it >> 1) & 1) ^ 1u);  // both epilogues drained this accumulator
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb, ++kbc) {
          const int s = kbc % STAGES;
          const uint32_t ph = (kbc / STAGES) & 1;
          mbar_wait(bar_fullA + 8 * s, ph);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t adesc = umma_desc_k_sw128(sA + s * G2_A_BYTES);
            const uint64_t bdesc = umma_desc_k_sw128(sB + s * G2_B_BYTES);
#pragma unroll
            for (int k = 0; k < G2_BK / 16; ++k)
              umma_f16_cg2(tbase + a * 256, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_cg2_mc(bar_empty + 8 * s, 3);
            if (kb == nkb - 1) umma_commit_cg2_mc(bar_tfull + 8 * a, 3);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 6) {
    // ================================ dequant warps ================================
    // (A variant in which each dequant warp took every 4th k-block passed the parity suite but could hang: with 5 stages and
    //  4 warps consecutive uses of a stage belong to different warps, and an mbarrier parity wait cannot tell use u from
    //  use u - 2 when the load of use u - 1 completes late — the same aliasing that faulted the small-batch tier; it also
    //  bought nothing, the k-block time here is not the dequant chain.  profiles/r02_midm_notes.md)
    const int t = threadIdx.x - 64;  // 0..127
    constexpr int ZSYM = 8;
    const int lp = t & 31, g = lp >> 2, tt = lp & 3;
    int f[4];
    f[0] = (t >> 5) * 16 + g;
    f[1] = f[0] + 8;
    f[2] = f[0] + 64;
    f[3] = f[0] + 72;
    const uint32_t bready_leader = mapa_u32(bar_fullA, 0);  // the leader's operands-ready barrier
    int kbc = 0;
    for (int tile = pair; tile < ntiles_total; tile += npairs) {
      for (int kb = 0; kb < nkb; ++kb, ++kbc) {
        const int s = kbc % STAGES;
        const uint32_t ph = (kbc / STAGES) & 1;
        mbar_wait(bar_fullP + 8 * s, ph);
        const uint8_t* pst = smem + (sP - smem_base) + s * G2_P_BYTES;
        const uint4* pj = reinterpret_cast<const uint4*>(pst);
        const int grow = ((2 * kb + (tt >> 1)) >> gshc) - ((2 * kb) >> gshc);  // 0 or 1
        const uint8_t* srow = pst + 4096 + grow * 320;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const uint4 pv = pj[t + u * 128];
          const uint32_t s_lo = *reinterpret_cast<const uint16_t*>(srow + f[2 * u] * 2);
          const uint32_t s_hi = *reinterpret_cast<const uint16_t*>(srow + f[2 * u + 1] * 2);
          int zl = ZSYM, zh = ZSYM;
          if (ASYM) {
            const uint32_t zwl = *reinterpret_cast<const uint32_t*>(srow + 256 + (f[2 * u] >> 3) * 4);
            const uint32_t zwh = *reinterpret_cast<const uint32_t*>(srow + 256 + (f[2 * u + 1] >> 3) * 4);
            zl = (int)((zwl >> (4 * g)) & 15u);
            zh = (int)((zwh >> (4 * g)) & 15u);
          }
          uint4 lo[2], hi[2];
          Dequant<T, 4>::run(pv, s_lo, zl, s_hi, zh, lo, hi);
          const uint32_t sw = (uint32_t)g;
          const uint32_t rlo = sB + s * G2_B_BYTES + f[2 * u] * 128;
          const uint32_t rhi = rlo + 8 * 128;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t off = (((uint32_t)(2 * tt + c)) ^ sw) << 4;
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(rlo + off), "r"(lo[c].x), "r"(lo[c].y),
                         "r"(lo[c].z), "r"(lo[c].w)
                         : "memory");
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(rhi + off), "r"(hi[c].x), "r"(hi[c].y),
                         "r"(hi[c].z), "r"(hi[c].w)
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (t == 0) mbar_arrive_cluster(bready_leader + 8 * s);
      }
    }
  } else {
    // ================================ epilogue warps: own 128 tokens x 256 features of every tile ============
    const int q = warp & 3;  // TMEM lane quarter
    const uint32_t tempty_leader = mapa_u32(bar_tempty, 0);
    int it = 0;
    for (int tile = pair; tile < ntiles_total; tile += npairs, ++it) {
      const int a = it & 1;
      const int tm = tile % TM;
      const G2Tile gt = g2_resolve(S, tile / TM);
      const int N = S.N[gt.set];
      const T* bias = reinterpret_cast<const T*>(S.bias[gt.set]);
      T* out = reinterpret_cast<T*>(S.out[gt.set]);
      const int npair0 = gt.tn * 256, m0 = tm * 256 + (int)rank * 128;
      mbar_wait(bar_tfull + 8 * a, (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      const int row = m0 + q * 32 + lane;
#pragma unroll 1
      for (int cc = 0; cc < 8; ++cc) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tbase + ((uint32_t)(q * 32) << 16) + a * 256 + cc * 32, r);
        tmem_ld_wait();
        const int nc = npair0 + cc * 32;
        if (row < M && nc < N) {
          T* dst = out + (size_t)row * N + nc;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float f0 = __uint_as_float(r[v * 8 + 2 * i]), f1 = __uint_as_float(r[v * 8 + 2 * i + 1]);
              if (bias != nullptr) {
                f0 = E::to_f(E::from_f(f0)) + E::to_f(bias[nc + v * 8 + 2 * i]);
                f1 = E::to_f(E::from_f(f1)) + E::to_f(bias[nc + v * 8 + 2 * i + 1]);
              }
              pk[i] = E::pack2(f0, f1);
            }
            *reinterpret_cast<uint4*>(dst + v * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      // accumulator a of this CTA is drained: one aggregated (possibly remote) arrive on the leader's barrier
      tc_fence_before();
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (threadIdx.x == 192) mbar_arrive_cluster(tempty_leader + 8 * a);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(G2P_TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
int make_x_tmap_box(CUtensorMap* map, const void* x, int M, int K, int dtype, int box_rows);  // b2q_gemm.cu (cached)

int make_x_tmap2(CUtensorMap* map, const void* x, int M, int K, int dtype) {
  return make_x_tmap_box(map, x, M, K, dtype, 128);
}

template <typename T, bool ASYM, int DQW>
static int launch_gemm2_t(const MmArgs& a, const void* x) {
  CUtensorMap tmap;
  if (make_x_tmap2(&tmap, x, a.M, a.K, a.dtype) != 0) return -1;
  auto kern = gemm2_kernel<T, ASYM, DQW>;
  static uint32_t smem_ok = 0;
  if (int e = ensure_dyn_smem(kern, G2_SMEM_BYTES, smem_ok, "b2q_gemm2")) return e;
  dim3 grid(2 * ((a.N + 255) / 256), (a.M + 255) / 256, 1);
  kern<<<grid, 64 + DQW * 32, G2_SMEM_BYTES, a.stream>>>(tmap, (const uint4*)a.packed, (const T*)a.scales,
                                                      (const uint32_t*)a.qzeros, (const T*)a.bias, (T*)a.out, a.M,
                                                      a.K, a.N, a.group_size, gemm_gshc(a));
  return (int)cudaGetLastError();
}

template <typename T, bool ASYM>
static int launch_gemm2p_t(const MmArgs& a, const void* x, const G2Sets& S) {
  CUtensorMap tmap;
  if (make_x_tmap2(&tmap, x, a.M, a.K, a.dtype) != 0) return -1;
  auto kern = gemm2p_kernel<T, ASYM>;
  static uint32_t smem_ok = 0;
  if (int e = ensure_dyn_smem(kern, G2_SMEM_BYTES, smem_ok, "b2q_gemm2p")) return e;
  const int TM = (a.M + 255) / 256, TN = S.tn_end[S.nsets - 1];
  const int tiles = TM * TN;
  const int npairs = tiles < 74 ? tiles : 74;
  kern<<<dim3(2 * npairs, 1, 1), G2P_THREADS, G2_SMEM_BYTES, a.stream>>>(tmap, S, a.M, a.K, a.group_size, gemm_gshc(a),
                                                                         TM, tiles);
  return (int)cudaGetLastError();
}

static int launch_gemm2p_sets(const MmArgs& a, const void* x, const G2Sets& S) {
  const bool asym = S.qzeros[0] != nullptr;
  if (a.dtype == 0) return asym ? launch_gemm2p_t<__half, true>(a, x, S) : launch_gemm2p_t<__half, false>(a, x, S);
  return asym ? launch_gemm2p_t<__nv_bfloat16, true>(a, x, S) : launch_gemm2p_t<__nv_bfloat16, false>(a, x, S);
}

// Sibling QuantLinears in one persistent launch (b2q_gemm_multi); x = activations with act-order already applied
int launch_gemm2_multi(const MmArgs& a, const void* x, int nsets, const void* const* packed, const void* const* scales,
                       const int32_t* const* qzeros, const void* const* bias, void* const* out, const int* Ns) {
  if (nsets < 1 || nsets > G2_MAX_SETS) {
    set_error("b2q_gemm_multi: nsets=%d out of range (1..%d)", nsets, G2_MAX_SETS);
    return -1;
  }
  G2Sets S = {};
  S.nsets = nsets;
  int tn = 0;
  for (int i = 0; i < nsets; ++i) {
    if (Ns[i] <= 0 || Ns[i] % 32 != 0 || packed[i] == nullptr || scales[i] == nullptr || out[i] == nullptr ||
        ((qzeros[i] != nullptr) != (qzeros[0] != nullptr))) {
      set_error("b2q_gemm_multi: set %d unsupported (N=%d; all sets share K, group size and symmetry)", i, Ns[i]);
      return -1;
    }
    tn += (Ns[i] + 255) / 256;
    S.tn_end[i] = tn;
    S.N[i] = Ns[i];
    S.packed[i] = (const uint4*)packed[i];
    S.scales[i] = scales[i];
    S.qzeros[i] = (const uint32_t*)qzeros[i];
    S.bias[i] = bias[i];
    S.out[i] = out[i];
  }
  for (int i = nsets; i < G2_MAX_SETS; ++i) S.tn_end[i] = tn;
  return launch_gemm2p_sets(a, x, S);
}

// x must already be the (act-order permuted, if any) activation matrix
int launch_gemm2(const MmArgs& a, const void* x) {
  const bool asym = a.qzeros != nullptr;
  if (env().gemm2_persist) {
    G2Sets S = {};
    S.nsets = 1;
    for (int i = 0; i < G2_MAX_SETS; ++i) S.tn_end[i] = (a.N + 255) / 256;
    S.N[0] = a.N;
    S.packed[0] = (const uint4*)a.packed;
    S.scales[0] = a.scales;
    S.qzeros[0] = (const uint32_t*)a.qzeros;
    S.bias[0] = a.bias;
    S.out[0] = a.out;
    return launch_gemm2p_sets(a, x, S);
  }
  const int dqw = env().gemm2_dqw;
#define B2Q_G2(T, AS) (dqw == 8 ? launch_gemm2_t<T, AS, 8>(a, x) : launch_gemm2_t<T, AS, 4>(a, x))
  if (a.dtype == 0) return asym ? B2Q_G2(__half, true) : B2Q_G2(__half, false);
  return asym ? B2Q_G2(__nv_bfloat16, true) : B2Q_G2(__nv_bfloat16, false);
#undef B2Q_G2
}

}  // namespace b2q
