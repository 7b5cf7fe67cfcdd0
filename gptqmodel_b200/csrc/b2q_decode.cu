// b2q_decode.cu — decode tier for 4-bit weights: out[M, N] = x[M, K] @ dequant(W) for M <= 8 (batch-1 decode and
// small speculative / multi-sequence batches), HBM-bound.
//
// 8,732,672 algorithmic bytes for 4096x4096 g128 at M=1; the roofline is the HBM copy bandwidth.
// The first version of this tier multiplied on the CUDA cores (fma.rn.f32.f16, kept in b2q_gemv.cu for 8-bit); ncu
// showed it issue-bound (156 warp-instructions per 1024 weights, profiles/r01_gemv_fhfma_v1.txt).  This version
// needs ~27:
//  * the prepacked T4 layout stores every 32-bit word as the four A-operand registers of one
//    mma.sync.m16n8k16 (16 features x 16 k): 4 LOP3 + 1 SHF produce them as exact 1024+q / 1024+16q halves, with
//    NO per-weight scaling, zero-point subtraction or conversion;
//  * the tensor pipe accumulates sum_k (bias+q) * x in fp32 for all (<= 8) tokens at once; once per quantisation
//    group the accumulators are folded into the running output with ONE fp32 fix-up per (feature, token):
//        out += s * (acc * c - (bias + z) * sum_k x)     (c = 1 or 1/16; sum_k x pre-reduced in shared memory)
//    which is algebraically the reference's  s * (q - z)  applied inside the sum (qlinear/__init__.py:1001-1003);
//  * grid = (N/32 feature tiles) x (KS split-K CTAs in a thread-block cluster); partial sums are reduced through
//    distributed shared memory (no atomics, no workspace, no output zeroing, deterministic);
//  * every lane issues its 4 x LDG.128 of weights (+ scales) BEFORE griddepcontrol.wait (programmatic dependent
//    launch): weights never depend on the previous kernel, so consecutive layers overlap their HBM streams;
//  * act-order: rows sorted by group at prepack, the x[perm[k']] gather is fused into the activation staging.
// Replaces the decode tiers of swordfish_mm (swordfish_mm.cu:216-286, mma.sync + cp.async + atomics) and Marlin's
// small-M path (marlin_template.h) in the reference.
#include "b2q_common.cuh"
#include "b2q_internal.h"

namespace b2q {

constexpr int DEC_MAX_WARPS = 8;
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

// One "quad" = 128 k x 32 features = 4 coalesced 512-byte rows of T4 (2 k-blocks of 64 x 2 feature tiles of 16),
// plus the scale / zero words of the (up to 2) groups it touches, for the 4 feature rows this lane owns.
template <bool ASYM, bool G64>
struct DQuad {
  uint4 v[4];                  // [kbl * 2 + ftl]
  uint16_t s[G64 ? 2 : 1][4];  // [group in quad][ftl * 2 + hi]
  uint32_t zw[(ASYM ? 1 : 0) * (G64 ? 2 : 1) + (ASYM ? 0 : 1)][4];
};

template <typename T, bool ASYM, bool G64>
__device__ __forceinline__ void load_dquad(DQuad<ASYM, G64>& q, const uint4* __restrict__ wp, size_t kb_stride,
                                           const T* __restrict__ scales, const uint32_t* __restrict__ qzeros,
                                           int quad, int gsh, int N, int nrow0) {
  // wp already points at (kb = 2*quad, ft = 2*nt, lane)
  q.v[0] = ldg_nc_v4(wp);
  q.v[1] = ldg_nc_v4(wp + 32);
  q.v[2] = ldg_nc_v4(wp + kb_stride);
  q.v[3] = ldg_nc_v4(wp + kb_stride + 32);
  const int g0 = (quad * 2) >> gsh;  // group of k-block 2*quad   (gsh = log2(group_size / 64), 31 for per-channel)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = nrow0 + (i >> 1) * 16 + (i & 1) * 8;
    q.s[0][i] = *reinterpret_cast<const uint16_t*>(scales + (size_t)g0 * N + n);
    if (ASYM) q.zw[0][i] = qzeros[(size_t)g0 * (N >> 3) + (n >> 3)];
  }
  if (G64) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = nrow0 + (i >> 1) * 16 + (i & 1) * 8;
      q.s[G64 ? 1 : 0][i] = *reinterpret_cast<const uint16_t*>(scales + (size_t)(g0 + 1) * N + n);
      if (ASYM) q.zw[G64 ? 1 : 0][i] = qzeros[(size_t)(g0 + 1) * (N >> 3) + (n >> 3)];
    }
  }
}

template <typename T, bool ASYM, bool G64>
__global__ void __launch_bounds__(DEC_MAX_WARPS * 32)
    decode_kernel(const uint4* __restrict__ packed, const T* __restrict__ scales, const uint32_t* __restrict__ qzeros,
                  const int32_t* __restrict__ perm, const T* __restrict__ x, const T* __restrict__ bias,
                  T* __restrict__ out, int M, int K, int N, int gsh, int qpc) {
  using E = ET<T>;
  extern __shared__ __align__(16) uint8_t dsm[];
  // dynamic smem: sx[M][kspan] (T) | xsum[kblocks][8] (float) | red[nwarps][8 acc][32] | part[8][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int nt = blockIdx.x, FT = N >> 4;
  const int nquads = K >> 7;
  const int q0 = blockIdx.y * qpc;
  const int q1 = min(q0 + qpc, nquads);
  const int kspan = qpc * 128;
  T* sx = reinterpret_cast<T*>(dsm);
  float* xsum = reinterpret_cast<float*>(dsm + (size_t)M * kspan * sizeof(T));
  float* red = xsum + qpc * 2 * 8;
  float* part = red + nwarps * 8 * 32;
  const bool PERM = perm != nullptr;
  const int nrow0 = nt * 32 + g;       // this lane's first feature row (others: +8, +16, +24)
  const size_t kb_stride = (size_t)FT * 32;

  // ---- 1. first quad of weights (+ scales) in flight before anything else ----------------------
  int q = q0 + warp;
  DQuad<ASYM, G64> cur;
  const uint4* wbase = packed + ((size_t)nt * 2) * 32 + lane;
  if (q < q1)
    load_dquad<T, ASYM, G64>(cur, wbase + (size_t)(2 * q) * kb_stride, kb_stride, scales, qzeros, q, gsh, N, nrow0);

  // PDL: let the next kernel start its own weight prefetch; wait for the producer of x only now.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  // ---- 2. stage x[m, k-range] (act-order gather fused) + per-(64 k block, token) sums ----------
  {
    const int n8 = (q1 - q0) * 16;   // uint4 (8 halves) per token row in this CTA's k-range
    const int tot = M * n8;
    const int totr = (tot + 31) & ~31;
    for (int i = threadIdx.x; i < totr; i += blockDim.x) {
      uint4 xv = make_uint4(0, 0, 0, 0);
      int m = 0, j = 0;
      if (i < tot) {
        m = i / n8;
        j = i - m * n8;
        const T* xr = x + (size_t)m * K;
        if (PERM) {
          const int4* pp = reinterpret_cast<const int4*>(perm + (size_t)q0 * 128) + 2 * j;
          const int4 p0 = pp[0], p1 = pp[1];
          const uint16_t* xu = reinterpret_cast<const uint16_t*>(xr);
          xv.x = (uint32_t)xu[p0.x] | ((uint32_t)xu[p0.y] << 16);
          xv.y = (uint32_t)xu[p0.z] | ((uint32_t)xu[p0.w] << 16);
          xv.z = (uint32_t)xu[p1.x] | ((uint32_t)xu[p1.y] << 16);
          xv.w = (uint32_t)xu[p1.z] | ((uint32_t)xu[p1.w] << 16);
        } else {
          xv = reinterpret_cast<const uint4*>(xr + (size_t)q0 * 128)[j];
        }
        reinterpret_cast<uint4*>(sx + (size_t)m * kspan)[j] = xv;
      }
      auto f2 = [](uint32_t u) {
        const T* h = reinterpret_cast<const T*>(&u);
        return E::to_f(h[0]) + E::to_f(h[1]);
      };
      float sm = (f2(xv.x) + f2(xv.y)) + (f2(xv.z) + f2(xv.w));
      sm += __shfl_xor_sync(0xffffffffu, sm, 1);
      sm += __shfl_xor_sync(0xffffffffu, sm, 2);
      sm += __shfl_xor_sync(0xffffffffu, sm, 4);
      if ((i & 7) == 0 && i < tot) xsum[(j >> 3) * 8 + m] = sm;  // 8 uint4 = one 64-k block
    }
    // zero the token columns >= M once (read by the fix-up of lanes whose columns are padding)
    for (int i = threadIdx.x; i < (q1 - q0) * 2 * 8; i += blockDim.x)
      if ((i & 7) >= M) xsum[i] = 0.f;
  }
  __syncthreads();

  // ---- 3. main loop ----------------------------------------------------------------------------
  constexpr float ZSYM = 8.f;
  float tot[2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) tot[a][b] = 0.f;

  while (q < q1) {
    const int qn = q + nwarps;
    DQuad<ASYM, G64> nxt;
    if (qn < q1)
      load_dquad<T, ASYM, G64>(nxt, wbase + (size_t)(2 * qn) * kb_stride, kb_stride, scales, qzeros, qn, gsh, N,
                               nrow0);
    float d[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) d[a][b] = 0.f;
    float xs0 = 0.f, xs1 = 0.f;  // sum_k x for token columns 2t, 2t+1 over the current group
#pragma unroll
    for (int kbl = 0; kbl < 2; ++kbl) {
      // activation fragment: token (column) g, k = 64*kb + 16t .. +15  -> 8 registers, 2 per k-step
      uint32_t bx[8];
      if (g < M) {
        const uint4* xp = reinterpret_cast<const uint4*>(sx + (size_t)g * kspan + ((q - q0) * 2 + kbl) * 64 + t * 16);
        const uint4 x0 = xp[0], x1 = xp[1];
        bx[0] = x0.x; bx[1] = x0.y; bx[2] = x0.z; bx[3] = x0.w;
        bx[4] = x1.x; bx[5] = x1.y; bx[6] = x1.z; bx[7] = x1.w;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) bx[i] = 0u;
      }
#pragma unroll
      for (int ftl = 0; ftl < 2; ++ftl) {
        const uint4 wv = cur.v[kbl * 2 + ftl];
        const uint32_t w[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          uint32_t a[4];
          E::unpack_w4(w[s], a);
          mma_16816<T>(d[ftl], a, bx[2 * s], bx[2 * s + 1]);
        }
      }
      const float2 xs = *reinterpret_cast<const float2*>(xsum + ((q - q0) * 2 + kbl) * 8 + 2 * t);
      xs0 += xs.x;
      xs1 += xs.y;
      if (kbl == 1 || G64) {
        // group boundary: fold the raw accumulators into the output with the per-(group, feature) scale / zero
        constexpr int dummy = 0;
        (void)dummy;
        const int gi = G64 ? kbl : 0;  // compile-time after unrolling
#pragma unroll
        for (int ftl = 0; ftl < 2; ++ftl) {
          const uint16_t slr = cur.s[gi][ftl * 2], shr = cur.s[gi][ftl * 2 + 1];
          const float sl = E::to_f(*reinterpret_cast<const T*>(&slr));
          const float sh = E::to_f(*reinterpret_cast<const T*>(&shr));
          float zl = ZSYM, zh = ZSYM;
          if (ASYM) {
            zl = (float)((cur.zw[ASYM ? gi : 0][ftl * 2] >> (4 * g)) & 15u);  // feature % 8 == g for all four rows
            zh = (float)((cur.zw[ASYM ? gi : 0][ftl * 2 + 1] >> (4 * g)) & 15u);
          }
          const float bl = E::LO_BASE + zl, bh = E::HI_BASE + zh;
          tot[ftl][0] = fmaf(sl, d[ftl][0] - bl * xs0, tot[ftl][0]);
          tot[ftl][1] = fmaf(sl, d[ftl][1] - bl * xs1, tot[ftl][1]);
          tot[ftl][2] = fmaf(sh, d[ftl][2] * E::HI_SCALE - bh * xs0, tot[ftl][2]);
          tot[ftl][3] = fmaf(sh, d[ftl][3] * E::HI_SCALE - bh * xs1, tot[ftl][3]);
          d[ftl][0] = d[ftl][1] = d[ftl][2] = d[ftl][3] = 0.f;
        }
        xs0 = xs1 = 0.f;
      }
    }
    if (qn < q1) cur = nxt;
    q = qn;
  }

  // ---- 4. reduce: warps -> CTA (smem) -> cluster (DSMEM) -> global -------------------------------
  // tot[ftl][c]: feature nt*32 + ftl*16 + g (+8 if c >= 2), token 2t + (c & 1)
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) red[(warp * 8 + a * 4 + b) * 32 + lane] = tot[a][b];
  __syncthreads();
  const uint32_t nrank = cluster_nctarank();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    float v = 0.f;
    for (int w = 0; w < nwarps; ++w) v += red[w * 256 + i];
    part[i] = v;
  }
  if (nrank > 1) cluster_sync_all(); else __syncthreads();
  if (cluster_ctarank() == 0) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
      const int acc = i >> 5, ln = i & 31;
      const int m = 2 * (ln & 3) + (acc & 1);
      if (m < M) {
        float v = part[i];
        for (uint32_t r = 1; r < nrank; ++r) v += ld_dsmem_f32(smem_u32(&part[i]), r);
        const int n = nt * 32 + (acc >> 2) * 16 + (ln >> 2) + ((acc & 2) ? 8 : 0);
        // reference order: round the matmul to the output dtype, then add bias (qlinear/torch.py:337-342)
        T o = E::from_f(v);
        if (bias != nullptr) o = E::from_f(E::to_f(o) + E::to_f(bias[n]));
        out[(size_t)m * N + n] = o;
      }
    }
  }
  if (nrank > 1) cluster_sync_all();  // keep peers' smem alive until rank 0 has read it
}

template <typename T, bool ASYM, bool G64>
static int launch_decode_t(const MmArgs& a, int ks, int warps, int qpc) {
  const size_t smem = (size_t)a.M * qpc * 128 * 2 + (size_t)qpc * 2 * 8 * 4 + (size_t)warps * 8 * 32 * 4 + 8 * 32 * 4;
  auto kern = decode_kernel<T, ASYM, G64>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a.N / 32, ks, 1);
  cfg.blockDim = dim3(warps * 32, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = a.stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = ks;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = a.pdl ? 2 : 1;
  if (ks > 8) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return (int)e;
  }
  int gsh = 31;  // per-channel: every k-block is group 0
  if (a.group_size == 64) gsh = 0;
  else if (a.group_size == 128) gsh = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, (const uint4*)a.packed, (const T*)a.scales,
                                     (const uint32_t*)a.qzeros, a.perm, (const T*)a.x, (const T*)a.bias, (T*)a.out,
                                     a.M, a.K, a.N, gsh, qpc);
  return (int)e;
}

bool decode_supported(const MmArgs& a) {
  return a.bits == 4 && a.M >= 1 && a.M <= DEC_MAXM && a.K % 128 == 0 && a.N % 32 == 0 &&
         (a.group_size == 64 || a.group_size == 128 || a.group_size == a.K);
}

int launch_decode(const MmArgs& a) {
  if (!decode_supported(a)) {
    set_error("b2q_decode: unsupported (bits=%d M=%d K=%d N=%d group=%d)", a.bits, a.M, a.K, a.N, a.group_size);
    return -1;
  }
  const int quads = a.K / 128, NT = a.N / 32;
  int warps = a.tune_warps > 0 ? a.tune_warps : 4;
  int ks;
  if (a.tune_ks > 0) {
    ks = a.tune_ks;
  } else {
    ks = 1;
    while (ks < 8 && NT * ks < 148 * 6 && quads / (ks * 2) >= warps) ks *= 2;
  }
  if (ks > quads) ks = quads;
  int qpc = (quads + ks - 1) / ks;
  // dynamic smem budget: M * qpc * 256 B of staged activations
  while ((size_t)a.M * qpc * 256 > 160 * 1024 && ks < 16) {
    ks *= 2;
    qpc = (quads + ks - 1) / ks;
  }
  if ((size_t)a.M * qpc * 256 > 160 * 1024) {
    set_error("b2q_decode: K=%d too large for M=%d", a.K, a.M);
    return -1;
  }
  const bool asym = a.qzeros != nullptr, g64 = a.group_size == 64;
#define B2Q_DEC_CASE(T)                                                                  \
  (asym ? (g64 ? launch_decode_t<T, true, true>(a, ks, warps, qpc)                       \
               : launch_decode_t<T, true, false>(a, ks, warps, qpc))                     \
        : (g64 ? launch_decode_t<T, false, true>(a, ks, warps, qpc)                      \
               : launch_decode_t<T, false, false>(a, ks, warps, qpc)))
  return a.dtype == 0 ? B2Q_DEC_CASE(__half) : B2Q_DEC_CASE(__nv_bfloat16);
#undef B2Q_DEC_CASE
}

}  // namespace b2q
