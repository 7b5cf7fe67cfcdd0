// b2q_gemm_sk.cu — single-CTA tcgen05 tier with CLUSTER SPLIT-K for small token counts (M <= 128).  EXPERIMENTAL:
// selected only with B2Q_GEMM_SPLITK=1 until it has passed the GPU parity suite.
//
// The plain single-CTA tier (b2q_gemm.cu) launches N/128 CTAs: 32 of 148 SMs for a 4096-wide layer, 28 us for
// 4096 x 4096 at any M <= 128 although the layer's 8.7 MB stream from HBM in 1.3 us (profiles/r01_gemm_notes.md).  Here
// `ks` CTAs of a thread-block cluster share one 128-feature tile: each runs the SAME warp-specialised pipeline (TMA
// activations, bulk-copied packed weights, exact dequant warps, tcgen05.mma into TMEM) over its own 1/ks of the
// k-blocks, parks its fp32 accumulator in its own shared memory, and after one cluster barrier every CTA reduces an
// interleaved share of the token rows over distributed shared memory (no atomics, no workspace, deterministic).
// 4096 x 4096: 32 tiles x ks = 4 -> 128 CTAs, 16 k-blocks each.
#include <cuda.h>

#include <cstdlib>

#include "b2q_common.cuh"
#include "b2q_dequant.cuh"
#include "b2q_gemm.cuh"
#include "b2q_internal.h"

namespace b2q {

template <typename T, int BITS, bool ASYM, int STAGES>
__global__ void __launch_bounds__(G_THREADS, 1)
    gemm_sk_kernel(const __grid_constant__ CUtensorMap tmap_x, const uint4* __restrict__ packed,
                const T* __restrict__ scales, const uint32_t* __restrict__ qzeros, const T* __restrict__ bias,
                T* __restrict__ out, int M, int K, int N, int group_size, int gshc, int kpc) {
  constexpr int MT = 1;
  using C = GemmCfg<BITS, MT, STAGES>;
  using E = ET<T>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t sA = smem_base;
  const uint32_t sB = sA + STAGES * C::A_BYTES;
  const uint32_t sP = sB + STAGES * C::B_BYTES;
  const uint32_t sBar = sP + STAGES * C::P_BYTES;
  const uint32_t bar_full = sBar, bar_bready = sBar + 8 * STAGES, bar_empty = sBar + 16 * STAGES;
  const uint32_t bar_tfull = sBar + 24 * STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + (sBar - smem_base) + 24 * STAGES + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NT = N >> 5;
  const int n0 = blockIdx.x * G_BN, m0 = blockIdx.y * (128 * MT);
  const int nt0 = n0 >> 5;
  const int ntiles = min(4, NT - nt0);
  // split-K: cluster rank z owns k-blocks [kb0, kb1); i = kb - kb0 drives the stage / phase bookkeeping
  const uint32_t nrank = cluster_nctarank(), crank = cluster_ctarank();
  const int kb0 = (int)crank * kpc, kb1 = min(K / G_BK, kb0 + kpc);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_bready + 8 * s, G_DQ_THREADS);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_ptr), C::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;

  if (warp == 0) {
    // ================================ producer ================================
    if (lane == 0) {
      // 4-bit: the 128 x 64 tile is one contiguous block of T4 (8 feature tiles of 16); 8-bit: two T8 rows
      const int FT = N >> 4, ft0 = n0 >> 4;
      const uint32_t pbytes4 = (uint32_t)min(8, FT - ft0) * 512u;
      const uint32_t pbytes8 = (uint32_t)ntiles * C::SUB * 512u;
      for (int kb = kb0; kb < kb1; ++kb) {
        const int s = (kb - kb0) % STAGES;
        const uint32_t ph = ((kb - kb0) / STAGES) & 1;
        mbar_wait(bar_empty + 8 * s, ph ^ 1);
        // 4-bit: scale / zero rows of the block's group(s) travel with the packed codes (no LDG in the dequant warps)
        const int g0 = (2 * kb) >> gshc, g1 = (2 * kb + 1) >> gshc;
        const int nrows = (g1 != g0) ? 2 : 1;
        const uint32_t sbytes = (uint32_t)min(128, N - n0) * 2u, zbytes = ASYM ? sbytes / 4u : 0u;
        mbar_expect_tx(bar_full + 8 * s,
                       C::A_BYTES + (BITS == 4 ? pbytes4 + nrows * (sbytes + zbytes) : 2 * pbytes8));
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
          tma_load_2d(sA + s * C::A_BYTES + mt * (128 * G_BK * 2), &tmap_x, bar_full + 8 * s, kb * G_BK,
                      m0 + mt * 128);
        if (BITS == 4) {
          bulk_load(sP + s * C::P_BYTES, packed + ((size_t)kb * FT + ft0) * 32, pbytes4, bar_full + 8 * s);
          for (int r = 0; r < nrows; ++r) {
            const int gr = r ? g1 : g0;
            bulk_load(sP + s * C::P_BYTES + 4096 + r * 320, scales + (size_t)gr * N + n0, sbytes, bar_full + 8 * s);
            if (ASYM)
              bulk_load(sP + s * C::P_BYTES + 4096 + r * 320 + 256, qzeros + (size_t)gr * (N >> 3) + (n0 >> 3),
                        zbytes, bar_full + 8 * s);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 2; ++j)
            bulk_load(sP + s * C::P_BYTES + j * C::P_CHUNK_BYTES,
                      packed + ((size_t)(kb * 2 + j) * NT + nt0) * C::SUB * 32, pbytes8, bar_full + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc = umma_idesc_f16(E::FMT, 128, G_BN);
    for (int kb = kb0; kb < kb1; ++kb) {
      const int s = (kb - kb0) % STAGES;
      const uint32_t ph = ((kb - kb0) / STAGES) & 1;
      mbar_wait(bar_full + 8 * s, ph);
      mbar_wait(bar_bready + 8 * s, ph);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t bdesc = umma_desc_k_sw128(sB + s * C::B_BYTES);
#pragma unroll
        for (int k = 0; k < G_BK / 16; ++k) {
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t adesc = umma_desc_k_sw128(sA + s * C::A_BYTES + mt * (128 * G_BK * 2));
            umma_f16(tbase + mt * G_BN, adesc + 2 * k, bdesc + 2 * k, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
        }
        umma_commit(bar_empty + 8 * s);
        if (kb == kb1 - 1) umma_commit(bar_tfull);
      }
      __syncwarp();
    }
  } else {
    // ================================ dequant warps ================================
    const int t = threadIdx.x - 64;  // 0..127
    constexpr int PF = 32 / BITS;
    constexpr int ZSYM = 1 << (BITS - 1);
    if (BITS == 4) {
      // thread owns two fragment-major uint4 per stage: feature tiles (t>>5) and (t>>5)+4, lane' = t&31 = 4g+tt
      const int lp = t & 31, g = lp >> 2, tt = lp & 3;
      int f[4];  // tile-local feature rows: lo/hi of the two uint4
      f[0] = (t >> 5) * 16 + g;
      f[1] = f[0] + 8;
      f[2] = f[0] + 64;
      f[3] = f[0] + 72;
      // this lane's 16 k of block kb lie in 32-k chunk 2*kb + (tt>>1); its scale / zero row was staged by the producer
      for (int kb = kb0; kb < kb1; ++kb) {
        const int s = (kb - kb0) % STAGES;
        const uint32_t ph = ((kb - kb0) / STAGES) & 1;
        mbar_wait(bar_full + 8 * s, ph);
        const uint8_t* pst = smem + (sP - smem_base) + s * C::P_BYTES;
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
            zl = (int)((zwl >> (4 * g)) & 15u);  // feature % 8 == g for both rows
            zh = (int)((zwh >> (4 * g)) & 15u);
          }
          uint4 lo[2], hi[2];
          Dequant<T, 4>::run(pv, s_lo, zl, s_hi, zh, lo, hi);
          const uint32_t sw = (uint32_t)g;  // (row & 7) for rows f and f+8
          const uint32_t rlo = sB + s * C::B_BYTES + f[2 * u] * 128;
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
        mbar_arrive(bar_bready + 8 * s);
      }
    } else {
      const int n = n0 + t;
      const int nsafe = (n < N) ? n : 0;
      const int ntl = t >> 5;
      SZRaw cur[2], nxt[2];
      cur[0] = load_sz<T, BITS, ASYM>(scales, qzeros, (2 * kb0) >> gshc, nsafe, N);
      cur[1] = load_sz<T, BITS, ASYM>(scales, qzeros, (2 * kb0 + 1) >> gshc, nsafe, N);
      for (int kb = kb0; kb < kb1; ++kb) {
        const int s = (kb - kb0) % STAGES;
        const uint32_t ph = ((kb - kb0) / STAGES) & 1;
        if (kb + 1 < kb1) {
          nxt[0] = load_sz<T, BITS, ASYM>(scales, qzeros, (2 * kb + 2) >> gshc, nsafe, N);
          nxt[1] = load_sz<T, BITS, ASYM>(scales, qzeros, (2 * kb + 3) >> gshc, nsafe, N);
        }
        mbar_wait(bar_full + 8 * s, ph);
        const uint32_t brow = sB + s * C::B_BYTES + t * 128;
        const uint32_t sw = (uint32_t)(t & 7);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          int z = ZSYM;
          if (ASYM) z = (int)((cur[j].zw >> (BITS * (nsafe % PF))) & ((1u << BITS) - 1));
          const uint4* pj = reinterpret_cast<const uint4*>(smem + (sP - smem_base) + s * C::P_BYTES +
                                                           j * C::P_CHUNK_BYTES);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint4 pv = pj[(ntl * 2 + h) * 32 + lane];
            uint4 o[2];
            Dequant<T, 8>::run(pv, cur[j].s, z, o);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const uint32_t addr = brow + (((uint32_t)(j * 4 + h * 2 + c) ^ sw) << 4);
              asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(o[c].x), "r"(o[c].y),
                           "r"(o[c].z), "r"(o[c].w)
                           : "memory");
            }
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(bar_bready + 8 * s);
        cur[0] = nxt[0];
        cur[1] = nxt[1];
      }
    }

    // ================================ epilogue ================================
    // 1. this rank's fp32 accumulator -> its own shared memory (the pipeline buffers are idle once bar_tfull fired:
    //    every TMA load was consumed by an MMA that has completed).  part[row][float4 chunk ^ (row & 31)]
    mbar_wait(bar_tfull, 0);
    tc_fence_after();
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    {
      const int prow = q * 32 + lane;
#pragma unroll
      for (int cc = 0; cc < G_BN / 32; ++cc) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tbase + ((uint32_t)(q * 32) << 16) + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          const uint32_t chunk = (uint32_t)(cc * 8 + v) ^ (uint32_t)(prow & 31);
          asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(sA + (uint32_t)prow * 512u + chunk * 16u),
                       "r"(r[4 * v]), "r"(r[4 * v + 1]), "r"(r[4 * v + 2]), "r"(r[4 * v + 3])
                       : "memory");
        }
      }
    }
    tc_fence_before();
  }
  // 2. all ranks' partials are in place (cluster barrier = CTA barrier + cross-CTA release/acquire)
  cluster_sync_all();
  if (warp >= 2) {
    // 3. rank z reduces token rows z, z + nrank, ... over all ranks through distributed shared memory
    const int t = threadIdx.x - 64;  // 0..127
    const int chunk = t & 31;        // float4 chunk (4 features) inside the 128-feature tile
    const int nc = n0 + chunk * 4;
    for (int rl = (int)crank + (int)nrank * (t >> 5); rl < 128; rl += (int)nrank * 4) {
      const int row = m0 + rl;
      if (row >= M || nc >= N) continue;
      const uint32_t local = sA + (uint32_t)rl * 512u + (((uint32_t)chunk ^ (uint32_t)(rl & 31)) << 4);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (uint32_t r = 0; r < nrank; ++r) {
        uint32_t ra;
        float4 v;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local), "r"(r));
        asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "r"(ra)
                     : "memory");
        acc[0] += v.x;
        acc[1] += v.y;
        acc[2] += v.z;
        acc[3] += v.w;
      }
      if (bias != nullptr) {
        // reference order: round the matmul to the output dtype, then add bias (torch.py:337-342)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = E::to_f(E::from_f(acc[i])) + E::to_f(bias[nc + i]);
      }
      *reinterpret_cast<uint2*>(out + (size_t)row * N + nc) = make_uint2(E::pack2(acc[0], acc[1]), E::pack2(acc[2], acc[3]));
    }
  }
  cluster_sync_all();  // keep every rank's shared memory alive until all peers have read it
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tbase, C::TMEM_COLS);
  }
}


int make_x_tmap(CUtensorMap* map, const void* x, int M, int K, int dtype);  // b2q_gemm.cu

// split-K ranks for an M <= 128 launch: fill ~148 SMs, keep >= 4 k-blocks per rank; 1 = not worth it
int gemm_sk_ranks(int K, int N) {
  const int tiles = (N + G_BN - 1) / G_BN, nkb = K / G_BK;
  int ks = 1;
  while (ks < 8 && tiles * ks * 2 <= 148 && nkb / (ks * 2) >= 4) ks *= 2;
  return ks;
}

template <typename T, int BITS, bool ASYM, int STAGES>
static int launch_gemm_sk_t(const MmArgs& a, const void* x, int ks) {
  using C = GemmCfg<BITS, 1, STAGES>;
  static_assert(STAGES * (C::A_BYTES + C::B_BYTES) >= 128 * 128 * 4, "the fp32 partial tile must fit the A/B stage buffers");
  CUtensorMap tmap;
  if (make_x_tmap(&tmap, x, a.M, a.K, a.dtype) != 0) return -1;
  auto kern = gemm_sk_kernel<T, BITS, ASYM, STAGES>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("b2q_gemm_sk: cannot opt in to %d bytes of shared memory: %s", C::SMEM_BYTES, cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  const int nkb = a.K / G_BK;
  const int kpc = (nkb + ks - 1) / ks;
  if ((ks - 1) * kpc >= nkb) {
    set_error("b2q_gemm_sk: ks=%d leaves a rank without k-blocks (K=%d)", ks, a.K);
    return -1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((a.N + G_BN - 1) / G_BN, 1, ks);
  cfg.blockDim = dim3(G_THREADS, 1, 1);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = a.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = ks;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmap, (const uint4*)a.packed, (const T*)a.scales,
                                     (const uint32_t*)a.qzeros, (const T*)a.bias, (T*)a.out, a.M, a.K, a.N,
                                     a.group_size, gemm_gshc(a), kpc);
  return (int)e;
}

// x must already be the (act-order permuted, if any) activation matrix; M <= 128
int launch_gemm_sk(const MmArgs& a, const void* x, int ks) {
  const bool asym = a.qzeros != nullptr;
#define B2Q_GSK(T, BITS, ST) (asym ? launch_gemm_sk_t<T, BITS, true, ST>(a, x, ks) : launch_gemm_sk_t<T, BITS, false, ST>(a, x, ks))
  if (a.dtype == 0) return a.bits == 4 ? B2Q_GSK(__half, 4, 4) : B2Q_GSK(__half, 8, 3);
  return a.bits == 4 ? B2Q_GSK(__nv_bfloat16, 4, 4) : B2Q_GSK(__nv_bfloat16, 8, 3);
#undef B2Q_GSK
}

}  // namespace b2q
