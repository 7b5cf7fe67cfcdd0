// b2q_gemm2s.cu — CTA-pair prefill kernel with a STREAM-K work split.  EXPERIMENTAL: selected only with
// B2Q_GEMM2_STREAMK=1 (and a zero-initialised workspace) until it has passed the GPU parity suite.
//
// Same pipeline as the persistent kernel of b2q_gemm2.cu (TMA producer | tcgen05.mma.cta_group::2 issuer | 4 dequant
// warps | 4 epilogue warps, two 256-column TMEM accumulators); what changes is the unit of work: not a whole 256 x 256
// tile but a (tile, k-block range) ITEM produced by SkIter (b2q_streamk.h).  The round-robin tile loop of the persistent
// kernel leaves the last wave partly empty (128 tiles on 74 pairs = 2 waves at 86 %, 448 tiles = 7 waves at 86 %:
// profiles/r01_gemm_notes.md); here the last R + P tiles are cut into P equal k-block streams, so all pairs finish
// together.  A pair whose segment starts inside a tile parks its fp32 accumulator in a workspace slot and bumps the
// tile's arrival counter; the pair that holds the tile's first k-block (at the END of its segment, i.e. later in time)
// waits for the counter, adds the parked partials in pair order (deterministic) and writes the tile.
//   workspace: u32 flags[2 * sk_tiles] (zero when first used; re-armed by the owner) | f32 slots[P][2][128][256]
#include <cuda.h>

#include <cstdlib>

#include "b2q_common.cuh"
#include "b2q_dequant.cuh"
#include "b2q_gemm2.cuh"
#include "b2q_internal.h"
#include "b2q_streamk.h"

namespace b2q {

constexpr int G2S_THREADS = 320;
constexpr int G2S_TMEM_COLS = 512;

__device__ __forceinline__ uint32_t sk_ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <typename T, bool ASYM>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2S_THREADS, 1)
    gemm2s_kernel(const __grid_constant__ CUtensorMap tmap_x, const uint4* __restrict__ packed,
                  const T* __restrict__ scales, const uint32_t* __restrict__ qzeros, const T* __restrict__ bias,
                  T* __restrict__ out, int M, int K, int N, int group_size, int gshc, int TM, const SkPlan pl,
                  uint32_t* __restrict__ sk_flags, float* __restrict__ sk_slots) {
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
  const int FT = N >> 4;
  const int pair = blockIdx.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_fullA + 8 * s, 1);
      mbar_init(bar_bready + 8 * s, 2);
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
                 "r"(G2S_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tbase = *tmem_ptr;

  if (warp == 0) {
    // ================================ producer ================================
    if (lane == 0) {
      int kbc = 0;  // k-block counter across work items (stage / phase bookkeeping)
      SkIter iter(pl, pair);
      SkItem it;
      while (iter.next(it)) {
        const int tm = it.tile % TM, tn = it.tile / TM;
        const int n0 = tn * 256 + (int)rank * 128, m0 = tm * 256 + (int)rank * 128;
        const int ft0 = n0 >> 4;
        const int nft = max(0, min(8, FT - ft0));
        const uint32_t pbytes = (uint32_t)nft * 512u;
        for (int kb = it.kb0; kb < it.kb1; ++kb, ++kbc) {
          const int s = kbc % STAGES;
          const uint32_t ph = (kbc / STAGES) & 1;
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
      int kbc = 0, n = 0;
      SkIter iter(pl, pair);
      SkItem it;
      for (; iter.next(it); ++n) {
        const int a = n & 1;
        mbar_wait(bar_tempty + 8 * a, (uint32_t)((n >> 1) & 1) ^ 1u);  // both epilogues drained this accumulator
        tc_fence_after();
        for (int kb = it.kb0; kb < it.kb1; ++kb, ++kbc) {
          const int s = kbc % STAGES;
          const uint32_t ph = (kbc / STAGES) & 1;
          mbar_wait(bar_fullA + 8 * s, ph);
          mbar_wait(bar_bready + 8 * s, ph);
          tc_fence_after();
          if (lane == 0) {
            const uint64_t adesc = umma_desc_k_sw128(sA + s * G2_A_BYTES);
            const uint64_t bdesc = umma_desc_k_sw128(sB + s * G2_B_BYTES);
#pragma unroll
            for (int k = 0; k < G2_BK / 16; ++k)
              umma_f16_cg2(tbase + a * 256, adesc + 2 * k, bdesc + 2 * k, idesc, (kb != it.kb0 || k != 0) ? 1u : 0u);
            umma_commit_cg2_mc(bar_empty + 8 * s, 3);
            if (kb == it.kb1 - 1) umma_commit_cg2_mc(bar_tfull + 8 * a, 3);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 6) {
    // ================================ dequant warps ================================
    const int t = threadIdx.x - 64;  // 0..127
    constexpr int ZSYM = 8;
    const int lp = t & 31, g = lp >> 2, tt = lp & 3;
    int f[4];
    f[0] = (t >> 5) * 16 + g;
    f[1] = f[0] + 8;
    f[2] = f[0] + 64;
    f[3] = f[0] + 72;
    const uint32_t bready_leader = mapa_u32(bar_bready, 0);
    int kbc = 0;
    SkIter iter(pl, pair);
    SkItem it;
    while (iter.next(it)) {
      for (int kb = it.kb0; kb < it.kb1; ++kb, ++kbc) {
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
    // ================================ epilogue warps: own 128 tokens x 256 features of every item ============
    const int q = warp & 3;  // TMEM lane quarter
    const uint32_t tempty_leader = mapa_u32(bar_tempty, 0);
    const int trow = q * 32 + lane;  // token row inside this CTA's 128-row half
    float* my_slot = sk_slots + ((size_t)pair * 2 + rank) * (size_t)(128 * 256);
    int n = 0;
    SkIter iter(pl, pair);
    SkItem it;
    for (; iter.next(it); ++n) {
      const int a = n & 1;
      const int tm = it.tile % TM, tn = it.tile / TM;
      const int npair0 = tn * 256, m0 = tm * 256 + (int)rank * 128;
      const bool first = it.kb0 == 0, last = it.kb1 == pl.nkb;
      const int row = m0 + trow;
      uint32_t* flag = sk_flags + (size_t)(it.tile - pl.dp_tiles) * 2 + rank;  // (split tiles only)
      // owner of a split tile: the pairs (pair, qlast] with a non-empty segment each deliver one partial
      int qlast = pair;
      if (first && !last) {
        qlast = sk_last_contributor(pl, it.tile);
        if (threadIdx.x == 192) {
          uint32_t expect = 0;
          for (int c = pair + 1; c <= qlast; ++c) expect += sk_begin(pl, c + 1) > sk_begin(pl, c) ? 1u : 0u;
          while (sk_ld_acquire_gpu(flag) < expect) {
          }
        }
      }
      mbar_wait(bar_tfull + 8 * a, (uint32_t)(n >> 1) & 1u);
      tc_fence_after();
      if (first && !last) asm volatile("bar.sync 2, 128;" ::: "memory");  // the partials are published
#pragma unroll 1
      for (int cc = 0; cc < 8; ++cc) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tbase + ((uint32_t)(q * 32) << 16) + a * 256 + cc * 32, r);
        tmem_ld_wait();
        const int nc = npair0 + cc * 32;
        if (row < M && nc < N) {
          if (!first) {
            // not the owner: park the fp32 partial accumulator (128 B per thread and column chunk)
            float4* dst = reinterpret_cast<float4*>(my_slot + (size_t)trow * 256 + cc * 32);
#pragma unroll
            for (int v = 0; v < 8; ++v)
              dst[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                   __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
          } else {
            for (int c = pair + 1; c <= qlast; ++c) {
              if (sk_begin(pl, c + 1) <= sk_begin(pl, c)) continue;  // empty segment
              const float4* ps = reinterpret_cast<const float4*>(sk_slots + ((size_t)c * 2 + rank) * (size_t)(128 * 256) +
                                                                 (size_t)trow * 256 + cc * 32);
#pragma unroll
              for (int v = 0; v < 8; ++v) {
                const float4 pv = __ldcg(ps + v);  // L2: the slot is rewritten by other SMs at every launch
                r[4 * v] = __float_as_uint(__uint_as_float(r[4 * v]) + pv.x);
                r[4 * v + 1] = __float_as_uint(__uint_as_float(r[4 * v + 1]) + pv.y);
                r[4 * v + 2] = __float_as_uint(__uint_as_float(r[4 * v + 2]) + pv.z);
                r[4 * v + 3] = __float_as_uint(__uint_as_float(r[4 * v + 3]) + pv.w);
              }
            }
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
      }
      if (!first) __threadfence();  // the partial is visible device-wide before the flag moves
      // accumulator a of this CTA is drained: one aggregated (possibly remote) arrive on the leader's barrier
      tc_fence_before();
      asm volatile("bar.sync 2, 128;" ::: "memory");
      if (threadIdx.x == 192) {
        mbar_arrive_cluster(tempty_leader + 8 * a);
        if (!first) atomicAdd(flag, 1u);              // contributor: one arrival per CTA
        else if (!last) *reinterpret_cast<volatile uint32_t*>(flag) = 0u;  // owner: all partials consumed, re-arm
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(G2S_TMEM_COLS) : "memory");
  }
}

int make_x_tmap2(CUtensorMap* map, const void* x, int M, int K, int dtype);  // b2q_gemm2.cu

size_t gemm2s_workspace_bytes() { return GEMM2S_FLAG_BYTES + (size_t)74 * 2 * 128 * 256 * sizeof(float); }

SkPlan gemm2s_plan(int M, int K, int N) {
  const int TM = (M + 255) / 256, TN = (N + 255) / 256;
  return sk_make_plan(TM * TN, K / G2_BK, 74);
}

// Host-side view of the work split (tests): plan5 = {tiles, P, nkb, dp_tiles, sk_tiles}; items = (tile, kb0, kb1, role)
// of `pair` in processing order, role 0 = whole tile, 1 = owner of a split tile (+ 16 * partials it waits for),
// 2 = contributor.  Returns the number of items (or -1 if more than max_items).
int gemm2s_debug_items(int M, int K, int N, int pair, int* plan5, int* items, int max_items) {
  const SkPlan pl = gemm2s_plan(M, K, N);
  plan5[0] = pl.tiles;
  plan5[1] = pl.P;
  plan5[2] = pl.nkb;
  plan5[3] = pl.dp_tiles;
  plan5[4] = pl.sk_tiles;
  if (pair < 0 || pair >= pl.P) return 0;
  SkIter iter(pl, pair);
  SkItem it;
  int n = 0;
  while (iter.next(it)) {
    if (n >= max_items) return -1;
    int role = 0;
    if (it.kb0 != 0) {
      role = 2;
    } else if (it.kb1 != pl.nkb) {
      const int qlast = sk_last_contributor(pl, it.tile);
      int expect = 0;
      for (int c = pair + 1; c <= qlast; ++c) expect += sk_begin(pl, c + 1) > sk_begin(pl, c) ? 1 : 0;
      role = 1 + 16 * expect;
    }
    items[4 * n] = it.tile;
    items[4 * n + 1] = it.kb0;
    items[4 * n + 2] = it.kb1;
    items[4 * n + 3] = role;
    ++n;
  }
  return n;
}

template <typename T, bool ASYM>
static int launch_gemm2s_t(const MmArgs& a, const void* x, void* ws) {
  CUtensorMap tmap;
  if (make_x_tmap2(&tmap, x, a.M, a.K, a.dtype) != 0) return -1;
  auto kern = gemm2s_kernel<T, ASYM>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES);
    if (e != cudaSuccess) {
      set_error("b2q_gemm2s: cannot opt in to %d bytes of shared memory: %s", G2_SMEM_BYTES, cudaGetErrorString(e));
      return (int)e;
    }
    attr_set = true;
  }
  const int TM = (a.M + 255) / 256;
  const SkPlan pl = gemm2s_plan(a.M, a.K, a.N);
  uint32_t* flags = reinterpret_cast<uint32_t*>(ws);
  float* slots = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ws) + GEMM2S_FLAG_BYTES);
  kern<<<dim3(2 * pl.P, 1, 1), G2S_THREADS, G2_SMEM_BYTES, a.stream>>>(
      tmap, (const uint4*)a.packed, (const T*)a.scales, (const uint32_t*)a.qzeros, (const T*)a.bias, (T*)a.out, a.M,
      a.K, a.N, a.group_size, gemm_gshc(a), TM, pl, flags, slots);
  return (int)cudaGetLastError();
}

// x must already be the (act-order permuted, if any) activation matrix; ws = gemm2s_workspace_bytes() bytes whose
// first GEMM2S_FLAG_BYTES were zero when the workspace was first used
int launch_gemm2s(const MmArgs& a, const void* x, void* ws) {
  const bool asym = a.qzeros != nullptr;
  if (a.dtype == 0) return asym ? launch_gemm2s_t<__half, true>(a, x, ws) : launch_gemm2s_t<__half, false>(a, x, ws);
  return asym ? launch_gemm2s_t<__nv_bfloat16, true>(a, x, ws) : launch_gemm2s_t<__nv_bfloat16, false>(a, x, ws);
}

}  // namespace b2q
