// b2q_gemm.cu — prefill / batched path: out[M, N] = x[M, K] @ dequant(W)[K, N] (+ bias) on tcgen05 tensor cores.
//
// One CTA computes a (128*MT tokens) x (128 features) output tile, K in blocks of 64, with a STAGES-deep
// mbarrier ring and warp-specialised roles:
//   warp 0      producer : TMA (cp.async.bulk.tensor, SWIZZLE_128B) for the activation tile and cp.async.bulk
//                          for the packed int4/int8 weight tile (contiguous 2 KB B2Q rows), one elected lane
//   warp 1      MMA      : allocates TMEM, one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//                          (M=128, N=128, K=16) with the fp32 accumulators in TMEM, tcgen05.commit -> mbarriers
//   warps 2..5  dequant  : LDS.128 packed weights -> exact (q - z) * s in fp16/bf16 (integer subtract first, one
//                          rounding: identical operands to the reference's torch dequant, qlinear/__init__.py:
//                          1001-1003) -> 16-byte swizzled K-major st.shared -> fence.proxy.async -> mbarrier;
//                          afterwards the same warps run the epilogue: tcgen05.ld TMEM -> regs -> (+bias) ->
//                          fp16/bf16 -> 64-byte-per-thread coalesced global stores.
// In the reference this is TorchLinear's dequant + torch.matmul (qlinear/torch.py:326-343), Marlin's
// mma.sync kernel (marlin_template.h) and Swordfish's CUTLASS-derived prefill tier (swordfish_prefill_*.cuh).
#include <cuda.h>

#include "b2q_common.cuh"
#include "b2q_dequant.cuh"
#include "b2q_gemm.cuh"
#include "b2q_internal.h"

namespace b2q {

template <typename T, int BITS, bool ASYM, int MT, int STAGES>
__global__ void __launch_bounds__(G_THREADS, 1)
    gemm_kernel(const __grid_constant__ CUtensorMap tmap_x, const uint4* __restrict__ packed,
                const T* __restrict__ scales, const uint32_t* __restrict__ qzeros, const T* __restrict__ bias,
                T* __restrict__ out, int M, int K, int N, int group_size, int gshc) {
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
  const int nkb = K / G_BK;

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
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
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
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
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
            umma_f16(tbase + mt * G_BN, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(bar_empty + 8 * s);
        if (kb == nkb - 1) umma_commit(bar_tfull);
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
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
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
      cur[0] = load_sz<T, BITS, ASYM>(scales, qzeros, 0, nsafe, N);
      cur[1] = load_sz<T, BITS, ASYM>(scales, qzeros, 1 >> gshc, nsafe, N);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        if (kb + 1 < nkb) {
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
    mbar_wait(bar_tfull, 0);
    tc_fence_after();
    const int q = warp & 3;  // TMEM lane quarter this warp may access
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int row = m0 + mt * 128 + q * 32 + lane;
#pragma unroll
      for (int cc = 0; cc < G_BN / 32; ++cc) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tbase + ((uint32_t)(q * 32) << 16) + mt * G_BN + cc * 32, r);
        tmem_ld_wait();
        const int nc = n0 + cc * 32;
        if (row < M && nc < N) {
          T* dst = out + (size_t)row * N + nc;
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              // reference order: round the matmul to the output dtype, then add bias (torch.py:337-342)
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
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tbase, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Tensor map of x [M, K] for boxes of (64 k) x (box_rows tokens), SWIZZLE_128B.  cuTensorMapEncodeTiled costs a few
// microseconds of host time: maps are cached per thread on (pointer, M, K, dtype, box) — the encoded descriptor depends
// on nothing else, so a hit is valid even if the allocation behind the pointer changed.
int make_x_tmap_box(CUtensorMap* map, const void* x, int M, int K, int dtype, int box_rows) {
  struct Entry {
    const void* x;
    int M, K, dtype, box;
    CUtensorMap map;
  };
  constexpr int NCACHE = 16;
  static thread_local Entry cache[NCACHE];
  static thread_local int next = 0, filled = 0;
  for (int i = 0; i < filled; ++i) {
    const Entry& c = cache[i];
    if (c.x == x && c.M == M && c.K == K && c.dtype == dtype && c.box == box_rows) {
      *map = c.map;
      return 0;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) {
    set_error("b2q_gemm: cuTensorMapEncodeTiled not available from the driver");
    return -1;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)M};
  cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)G_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                   const_cast<void*>(x), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("b2q_gemm: cuTensorMapEncodeTiled failed (%d) for x=%p M=%d K=%d box=%d", (int)r, x, M, K, box_rows);
    return -1;
  }
  Entry& e = cache[next];
  e.x = x;
  e.M = M;
  e.K = K;
  e.dtype = dtype;
  e.box = box_rows;
  e.map = *map;
  next = (next + 1) % NCACHE;
  if (filled < NCACHE) ++filled;
  return 0;
}

int make_x_tmap(CUtensorMap* map, const void* x, int M, int K, int dtype) {
  return make_x_tmap_box(map, x, M, K, dtype, 128);
}

// log2(32-k chunks per group); 31 for per-channel (every chunk maps to group 0)
int gemm_gshc(const MmArgs& a) {
  if (a.group_size == 32) return 0;
  if (a.group_size == 64) return 1;
  if (a.group_size == 128) return 2;
  return 31;
}

template <typename T, int BITS, bool ASYM, int MT, int STAGES>
static int launch_gemm_t(const MmArgs& a, const void* x) {
  using C = GemmCfg<BITS, MT, STAGES>;
  CUtensorMap tmap;
  if (make_x_tmap(&tmap, x, a.M, a.K, a.dtype) != 0) return -1;
  auto kern = gemm_kernel<T, BITS, ASYM, MT, STAGES>;
  static uint32_t smem_ok = 0;
  if (int e = ensure_dyn_smem(kern, C::SMEM_BYTES, smem_ok, "b2q_gemm")) return e;
  dim3 grid((a.N + G_BN - 1) / G_BN, (a.M + 128 * MT - 1) / (128 * MT), 1);
  kern<<<grid, G_THREADS, C::SMEM_BYTES, a.stream>>>(tmap, (const uint4*)a.packed, (const T*)a.scales,
                                                     (const uint32_t*)a.qzeros, (const T*)a.bias, (T*)a.out, a.M,
                                                     a.K, a.N, a.group_size, gemm_gshc(a));
  return (int)cudaGetLastError();
}

int launch_gemm(const MmArgs& a) {
  if (a.K % G_BK != 0) {
    set_error("b2q_gemm: K=%d must be a multiple of %d", a.K, G_BK);
    return -1;
  }
  const void* x = a.x;
  if (a.perm != nullptr) {
    const size_t need = (size_t)a.M * a.K * 2;
    if (a.workspace == nullptr || a.workspace_bytes < need) {
      set_error("b2q_gemm: act-order needs a %zu-byte workspace (got %zu)", need, a.workspace_bytes);
      return -1;
    }
    int e = launch_permute_cols(a.x, a.perm, a.workspace, a.M, a.K, a.stream);
    if (e != 0) return e;
    x = a.workspace;
  }
  // M <= 128: small-batch tier (swapped operands, cluster split-K); B2Q_MIDM=0 keeps the round-1 padded single-CTA path
  if (a.M <= 128 && env().midm && midm_supported(a)) return launch_midm(a, x);
  if (a.bits == 4 && a.M > 128 && a.tune_ks != -1)  // CTA-pair tier (tune_ks -1: force 1-CTA)
    return launch_gemm2(a, x);
  const bool asym = a.qzeros != nullptr;
  const bool big = a.M > 128;
#define B2Q_GEMM_CASE(T, BITS, ST)                                                              \
  (asym ? (big ? launch_gemm_t<T, BITS, true, 2, ST>(a, x) : launch_gemm_t<T, BITS, true, 1, ST>(a, x)) \
        : (big ? launch_gemm_t<T, BITS, false, 2, ST>(a, x) : launch_gemm_t<T, BITS, false, 1, ST>(a, x)))
  if (a.dtype == 0) return a.bits == 4 ? B2Q_GEMM_CASE(__half, 4, 4) : B2Q_GEMM_CASE(__half, 8, 3);
  return a.bits == 4 ? B2Q_GEMM_CASE(__nv_bfloat16, 4, 4) : B2Q_GEMM_CASE(__nv_bfloat16, 8, 3);
#undef B2Q_GEMM_CASE
}

}  // namespace b2q
