// b2q_midm.cu — small-batch tier (2 <= tokens <= 128; every 8-bit / group_size 32 shape the decode tier does not
// take): out[M, N] = x[M, K] @ dequant(W)[K, N] (+ bias) with SWAPPED tcgen05 operands and cluster split-K.
//
// Batched decode / speculative decoding is HBM-bound like batch-1 decode: the layer's packed weights must stream once,
// the token count only changes the width of the MMA.  The prefill tiers pad tokens to the UMMA M = 128 and launch
// N/128 CTAs (32 of 148 SMs for a 4096-wide layer: 28 us for 4096 x 4096 at any M <= 128, 0.05 of the HBM roofline —
// VERDICT r01 weak #4).  Here
//   * the WEIGHTS are the UMMA A operand: 128 output features = the M = 128 rows of tcgen05.mma.cta_group::1.kind::f16,
//     the TOKENS are the B operand, N = NTOK in {16, 32, 64, 128}: no padding of tokens to 128, the accumulator
//     D[feature][token] needs only NTOK TMEM columns and the x tile is NTOK x 64 k (TMA, SWIZZLE_128B);
//   * `ks` CTAs of a thread-block cluster split the k-blocks of one 128-feature tile (4096 x 4096: 32 tiles x 4 = 128
//     CTAs), park their fp32 partials TRANSPOSED ([token][feature]) in their own shared memory and reduce an interleaved
//     share of the token rows over distributed shared memory: no atomics, no workspace, deterministic, and the global
//     stores are 256 contiguous bytes per warp (features are the fast axis of `out`);
//   * 8 dequant warps (one fragment-major uint4 = 2 features x 16 k per thread and k-block) produce the exact
//     (q - z) * s operand (integer subtract first, ONE rounding: qlinear/__init__.py:1001-1003) as 16-byte swizzled
//     K-major rows.  The dequant warps form 4 groups that take every 4th k-block; each group's leader issues the bulk copies of
//     the group's own packed stages, warp 0 only the x tile TMA: no single thread issues every asynchronous copy of every
//     k-block, and the weight stream + dequant of the first blocks run BEFORE griddepcontrol.wait (programmatic dependent
//     launch), i.e. under the previous kernel's tail.
// Reference counterparts: Swordfish's Stream-K / atomic split-K decode for 17 <= M < 128 (swordfish_mm.cu:113-154,
// 229-276), its swapped-problem-shape tcgen05 trick (swordfish_prefill_impl.cuh:219-227), Marlin's small-M tiles
// (marlin_template.h).
#include <cuda.h>

#include "b2q_common.cuh"
#include "b2q_dequant.cuh"
#include "b2q_internal.h"

namespace b2q {

constexpr int MM_BF = 128;                  // features per tile (UMMA M)
constexpr int MM_BK = 64;                   // k per block (one SWIZZLE_128B row of fp16)
constexpr int MM_DQ_WARPS = 8;
constexpr int MM_DQ_THREADS = MM_DQ_WARPS * 32;
constexpr int MM_THREADS = 64 + MM_DQ_THREADS + 96;  // producer, MMA issuer 0, 8 dequant warps, MMA issuers 1..3 (warps 10..12)

// Two rings.  The PACKED ring (PST stages: x tile + packed codes + scale / zero rows, 7-21 KB each) is what hides the HBM
// latency: with 4 stages (the first version of this kernel) only 16 KB of weights per SM were in flight and every k-block
// cost ~0.39 us = one HBM round trip / 4 (profiles/r02_midm_v1_bench.log: 25 us for the 64 k-blocks of 4096 x 4096 on one
// CTA per tile); Little's law asks for ~43 KB per SM at 6.5 TB/s.  The DEQUANTISED ring (WST stages of 16 KB) only
// decouples the dequant warps from the tensor core.
// MODE 0: one QuantLinear.  MODE 1 / 2: GROUPED over the experts of a MoE block (b2q_moe.cu): blockIdx.z = (expert, token
// block); the rows of an expert are contiguous in the expert-sorted activation matrix; CTAs beyond an expert's row count
// exit at once.  MODE 1 runs TWO weight sets (w1 = gate, w3 = up) through the pipeline back to back into two TMEM
// accumulators and stores silu(gate) * up; MODE 2 (w2 = down) scales each row by its routing weight and scatters it to the
// row's (token, k) slot of an fp32 buffer.
struct MoeArgs {
  const int32_t* counts;        // [E] rows of expert e
  const int32_t* offsets;       // [E] first row of expert e in the sorted order
  const int32_t* sorted_pairs;  // [rows] pair index (token * top_k + j) of sorted row i            (MODE 2)
  const float* pair_weights;    // [rows] routing weight, indexed by PAIR index                      (MODE 2)
  const uint4* packed3;         // second weight set, same shapes / strides as the first             (MODE 1)
  const void* scales3;
  const uint32_t* qzeros3;
  float* ypair;                 // [rows, N] fp32, row = pair index                                   (MODE 2)
  int tblocks;                  // token blocks (of NTOK rows) per expert in gridDim.z
};

template <int BITS, int NTOK, int PST, int WST, int MODE = 0>
struct MidCfg {
  static constexpr int NSETS = MODE == 1 ? 2 : 1;
  static constexpr int XST = NTOK >= 64 ? 4 : 8;         // activation-tile stages (own ring: one TMA per k-block)
  static constexpr int SUB = BITS / 4;
  static constexpr int W_BYTES = MM_BF * MM_BK * 2;      // dequantised weights, A operand: 16 KB
  static constexpr int X_BYTES = NTOK * MM_BK * 2;       // activations, B operand (first in its stage: 1024-B aligned)
  static constexpr int P_CHUNK_BYTES = 4 * SUB * 512;    // 8-bit: 4 feature tiles (of 32) x 32 k
  static constexpr int P_BYTES = 2 * P_CHUNK_BYTES + (BITS == 4 ? 1024 : 0);  // + scale / zero rows (4-bit)
  static constexpr int BAR_BYTES = 512;
  static constexpr int RING_BYTES = WST * W_BYTES + XST * X_BYTES + PST * P_BYTES;
  static constexpr int SMEM_BYTES = RING_BYTES + BAR_BYTES + 1024;
  static constexpr int ACC_COLS = NSETS * NTOK;          // one accumulator set per MMA issuer
  static constexpr int NISS = 4 * ACC_COLS <= 512 ? 4 : 2;  // MMA issuer warps (k-blocks i, i + NISS, ... each)
  static constexpr int TMEM_COLS = NISS * ACC_COLS < 32 ? 32 : NISS * ACC_COLS;
  static_assert(XST % NISS == 0 && WST % NISS == 0, "every use of a stage must belong to the same MMA issuer");
  static constexpr int PART_BYTES = NSETS * NTOK * MM_BF * 4;  // fp32 partial tile(s) [set][token][feature]
  static_assert(MODE == 0 || BITS == 4, "the grouped (MoE) modes are built for 4-bit experts");
  static_assert(X_BYTES % 1024 == 0, "x tiles must stay 1024-byte aligned (SWIZZLE_128B atoms)");
  static_assert(PART_BYTES <= RING_BYTES, "the fp32 partial tile reuses the idle stage buffers");
  static_assert((PST + 2 * XST + 2 * WST + 1) * 8 + 16 <= BAR_BYTES, "mbarrier area");
  static_assert(SMEM_BYTES <= 227 * 1024, "dynamic shared memory of one CTA");
};

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

template <typename T, int BITS, bool ASYM, int NTOK, int PST, int WST, int MODE, int DQG>
__global__ void __launch_bounds__(MM_THREADS, 1)
    midm_kernel(const __grid_constant__ CUtensorMap tmap_x, const uint4* __restrict__ packed,
                const T* __restrict__ scales, const uint32_t* __restrict__ qzeros, const T* __restrict__ bias,
                T* __restrict__ out, int M, int K, int N, int gshc, int kpc, const __grid_constant__ MoeArgs G) {
  using C = MidCfg<BITS, NTOK, PST, WST, MODE>;
  using E = ET<T>;
  constexpr int NSETS = C::NSETS;
  int row0 = 0;  // first row of this CTA's token block in the activation matrix
  const uint4* packed_b = nullptr;
  const T* scales_b = nullptr;
  const uint32_t* qzeros_b = nullptr;
  if (MODE != 0) {
    // the routing tables are written by the preceding kernel of the stream: nothing may be read before it has finished
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int e = (int)blockIdx.z / G.tblocks, tb = (int)blockIdx.z - e * G.tblocks;
    const int cnt = G.counts[e];
    if (tb * NTOK >= cnt) return;  // same decision in every CTA of the cluster (they differ in blockIdx.y only)
    row0 = G.offsets[e] + tb * NTOK;
    M = min(NTOK, cnt - tb * NTOK);
    const size_t groups = (size_t)(K >> 5) >> gshc;  // quantisation groups along K (gshc = log2(32-k chunks per group))
    const size_t wstride = (size_t)K * N / 32, sstride = (gshc >= 31 ? 1 : groups) * (size_t)N;
    packed += (size_t)e * wstride;
    scales += (size_t)e * sstride;
    if (ASYM) qzeros += (size_t)e * (sstride >> 3);
    if (MODE == 1) {
      packed_b = G.packed3 + (size_t)e * wstride;
      scales_b = reinterpret_cast<const T*>(G.scales3) + (size_t)e * sstride;
      if (ASYM) qzeros_b = G.qzeros3 + (size_t)e * (sstride >> 3);
    }
  }
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

  constexpr int XST = C::XST;
  constexpr bool ISSUER_REFILL = PST >= 3 * DQG;  // who refills the packed stages (see the dequant warps)
  const uint32_t sW = smem_base;                      // [WST][128 features][64 k]      (also: fp32 partial tile)
  const uint32_t sXr = sW + WST * C::W_BYTES;         // [XST] x tile [NTOK][64 k]
  const uint32_t sPr = sXr + XST * C::X_BYTES;        // [PST]{ packed codes | scale / zero rows }
  const uint32_t sBar = sPr + PST * C::P_BYTES;
  const uint32_t bar_pfull = sBar, bar_xfull = bar_pfull + 8 * PST, bar_xempty = bar_xfull + 8 * XST;
  const uint32_t bar_wready = bar_xempty + 8 * XST, bar_wempty = bar_wready + 8 * WST, bar_tfull = bar_wempty + 8 * WST;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + (bar_tfull - smem_base) + 8);
  auto sXs = [&](int s) { return sXr + (uint32_t)s * C::X_BYTES; };
  auto sPs = [&](int s) { return sPr + (uint32_t)s * C::P_BYTES; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NT = N >> 5;
  const int n0 = blockIdx.x * MM_BF;
  const int nt0 = n0 >> 5;
  const int ntiles = min(4, NT - nt0);
  // split-K: cluster rank owns k-blocks [kb0, kb1); i = kb - kb0 drives the stage / phase bookkeeping
  const uint32_t nrank = cluster_nctarank(), crank = cluster_ctarank();
  const int kb0 = (int)crank * kpc, kb1 = min(K / MM_BK, kb0 + kpc);
  const int nkb = kb1 - kb0;
  const int NI = NSETS * nkb;  // pipeline iterations: the k-blocks of set 0, then (MODE 1) of set 1

  // PDL: the next kernel in the stream may start its own prologue / weight prefetch now
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmap_x);
    for (int s = 0; s < PST; ++s) mbar_init(bar_pfull + 8 * s, 1);
    for (int s = 0; s < XST; ++s) {
      mbar_init(bar_xfull + 8 * s, 1);
      mbar_init(bar_xempty + 8 * s, 1);
    }
    for (int s = 0; s < WST; ++s) {
      mbar_init(bar_wready + 8 * s, MM_DQ_THREADS / DQG);
      mbar_init(bar_wempty + 8 * s, 1);
    }
    mbar_init(bar_tfull, nkb >= C::NISS ? C::NISS : (nkb >= 2 ? 2 : 1));  // one arrival per active MMA issuer
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

  // ---- weight loads of pipeline iteration i into packed stage s (issued by the LEADER of the dequant group that will
  // consume the stage, see below: one thread per SM issuing every asynchronous copy of every k-block was the k-block time,
  // ~150 clk per UBLKCP / UTMALDG from a single thread — profiles/r02_midm_notes.md)
  const int FT = N >> 4, ft0 = n0 >> 4;
  const uint32_t pbytes4 = (uint32_t)min(8, FT - ft0) * 512u;
  const uint32_t pbytes8 = (uint32_t)ntiles * C::SUB * 512u;
  const uint32_t sbytes = (uint32_t)min(MM_BF, N - n0) * 2u, zbytes = ASYM ? sbytes / 4u : 0u;
  auto load_weights = [&](int i, int s) {
    const bool second = NSETS > 1 && i >= nkb;
    const int kb = kb0 + (second ? i - nkb : i);
    const uint4* pk = second ? packed_b : packed;
    const T* sc = second ? scales_b : scales;
    const uint32_t* zq = second ? qzeros_b : qzeros;
    if (BITS == 4) {
      // the scale / zero rows of the block's group(s) travel with the packed codes (no LDG in the dequant warps)
      const int g0 = (2 * kb) >> gshc, g1 = (2 * kb + 1) >> gshc;
      const int nrows = (g1 != g0) ? 2 : 1;
      mbar_expect_tx(bar_pfull + 8 * s, pbytes4 + nrows * (sbytes + zbytes));
      bulk_load(sPs(s), pk + ((size_t)kb * FT + ft0) * 32, pbytes4, bar_pfull + 8 * s);
      for (int r = 0; r < nrows; ++r) {
        const int gr = r ? g1 : g0;
        bulk_load(sPs(s) + 4096 + r * 320, sc + (size_t)gr * N + n0, sbytes, bar_pfull + 8 * s);
        if (ASYM)
          bulk_load(sPs(s) + 4096 + r * 320 + 256, zq + (size_t)gr * (N >> 3) + (n0 >> 3), zbytes, bar_pfull + 8 * s);
      }
    } else {
      mbar_expect_tx(bar_pfull + 8 * s, 2 * pbytes8);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        bulk_load(sPs(s) + j * C::P_CHUNK_BYTES, pk + ((size_t)(kb * 2 + j) * NT + nt0) * C::SUB * 32, pbytes8,
                  bar_pfull + 8 * s);
    }
  };

  if (warp == 0) {
    // ================================ activation producer ================================
    // one TMA per pipeline iteration: the x tile of the k-block (the weights are loaded by the dequant group leaders)
    if (lane == 0) {
      asm volatile("griddepcontrol.wait;" ::: "memory");  // x is the previous kernel's output
      for (int i = 0; i < NI; ++i) {
        const int xs = i % XST;
        if (i >= XST) mbar_wait(bar_xempty + 8 * xs, ((i / XST) & 1) ^ 1);
        const int kb = kb0 + ((NSETS > 1 && i >= nkb) ? i - nkb : i);
        mbar_expect_tx(bar_xfull + 8 * xs, C::X_BYTES);
        tma_load_2d(sXs(xs), &tmap_x, bar_xfull + 8 * xs, kb * MM_BK, row0);
      }
    }
  } else if (warp == 1 || warp >= 10) {
    // ================================ MMA issuers ================================
    // Up to FOUR issuing warps take every NISS-th pipeline iteration and accumulate into their OWN TMEM columns (summed in
    // the epilogue: deterministic).  One thread issuing the two mbarrier waits, four tcgen05.mma and two or three
    // tcgen05.commit of EVERY k-block was the k-block time of this tier (~770 clk at any token width, ring depth, dequant
    // parallelism or copy-issue scheme: profiles/r02_midm_notes.md).
    constexpr uint32_t idesc = umma_idesc_f16(E::FMT, MM_BF, NTOK);
    const int nissue = nkb >= C::NISS ? C::NISS : (nkb >= 2 ? 2 : 1);
    const int mi = warp == 1 ? 0 : warp - 9;
    for (int i = mi; i < NI && mi < nissue; i += nissue) {
      const int xs = i % XST, ws = i % WST;
      mbar_wait(bar_xfull + 8 * xs, (i / XST) & 1);
      mbar_wait(bar_wready + 8 * ws, (i / WST) & 1);
      tc_fence_after();
      if (lane == 0) {
        const bool second = NSETS > 1 && i >= nkb;
        // accumulator of (issuer mi, set): TMEM columns (mi * NSETS + set) * NTOK
        const uint32_t dcol = tbase + (uint32_t)(mi * C::ACC_COLS + (second ? NTOK : 0));
        const int start = second ? nkb : 0;
        const int first = start + ((mi - start % nissue + nissue) % nissue);  // this issuer's first block of the set
        const uint64_t wdesc = umma_desc_k_sw128(sW + ws * C::W_BYTES);
        const uint64_t xdesc = umma_desc_k_sw128(sXs(xs));
#pragma unroll
        for (int k = 0; k < MM_BK / 16; ++k)
          umma_f16(dcol, wdesc + 2 * k, xdesc + 2 * k, idesc, (i != first || k != 0) ? 1u : 0u);
        umma_commit(bar_wempty + 8 * ws);  // the dequantised stage may be overwritten
        umma_commit(bar_xempty + 8 * xs);  // the x tile has been read
        if (i + nissue >= NI) umma_commit(bar_tfull);  // this issuer's last block
        // wready(i) completed => the dequant group has read packed stage i % PST: stream block i + PST into it (one
        // issuer thread per k-block class, off the dequant groups' latency chain)
        if (ISSUER_REFILL && i + PST < NI) load_weights(i + PST, i % PST);
      }
      __syncwarp();
    }
  } else if (warp < 10) {
    // ================================ dequant warps ================================
    // The dequant warps form DQG groups of TG threads; group gq takes the pipeline iterations i = gq, gq + DQG, ...  One
    // iteration is a serial chain of latencies for a warp (mbarrier wake-up, LDS, ~75 ALU instructions per uint4, STS,
    // fence.proxy.async, arrive: ~800 clk) — with all eight warps on the SAME k-block (first version) that chain WAS the
    // k-block time (0.43 us per k-block at any ring depth, profiles/r02_midm_notes.md); with DQG blocks in flight the
    // chains overlap and the tier becomes issue-bound instead.
    const int t = threadIdx.x - 64;  // 0..255
    constexpr int TG = MM_DQ_THREADS / DQG;
    static_assert(PST % DQG == 0 && WST % DQG == 0, "every use of a ring stage must belong to the same dequant group");
    const int gq = t / TG, tl = t - gq * TG;
    constexpr int PF = 32 / BITS;
    constexpr int ZSYM = 1 << (BITS - 1);
    // the group's leader starts the weight stream of the group's first PST / DQG blocks (never depends on the previous
    // kernel: under programmatic dependent launch this runs while the producer of x is still executing)
    if (tl == 0)
      for (int i = gq; i < NI && i < PST; i += DQG) load_weights(i, i);
    if (BITS == 4) {
      // DQG fragment-major uint4 per thread and iteration: q = tl + TG*u -> feature tile q>>5 (16 features), lane' = q&31
      constexpr int U = DQG;
      const int lp = tl & 31, g = lp >> 2, tt = lp & 3;
      for (int i = gq; i < NI; i += DQG) {
        const int kb = kb0 + ((NSETS > 1 && i >= nkb) ? i - nkb : i), s = i % PST, ws = i % WST;
        mbar_wait(bar_pfull + 8 * s, (i / PST) & 1);
        const uint8_t* pst = smem + (sPs(s) - smem_base);
        // this lane's 16 k of block kb lie in 32-k chunk 2*kb + (tt>>1)
        const int grow = ((2 * kb + (tt >> 1)) >> gshc) - ((2 * kb) >> gshc);  // 0 or 1
        const uint8_t* srow = pst + 4096 + grow * 320;
        uint4 pv[U];
        uint32_t s_lo[U], s_hi[U];
        int zl[U], zh[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int q = tl + TG * u;
          const int f_lo = (q >> 5) * 16 + g, f_hi = f_lo + 8;
          pv[u] = reinterpret_cast<const uint4*>(pst)[q];
          s_lo[u] = *reinterpret_cast<const uint16_t*>(srow + f_lo * 2);
          s_hi[u] = *reinterpret_cast<const uint16_t*>(srow + f_hi * 2);
          zl[u] = zh[u] = ZSYM;
          if (ASYM) {
            const uint32_t zwl = *reinterpret_cast<const uint32_t*>(srow + 256 + (f_lo >> 3) * 4);
            const uint32_t zwh = *reinterpret_cast<const uint32_t*>(srow + 256 + (f_hi >> 3) * 4);
            zl[u] = (int)((zwl >> (4 * g)) & 15u);  // feature % 8 == g for both rows
            zh[u] = (int)((zwh >> (4 * g)) & 15u);
          }
        }
        // Refill of the packed stage for block i + PST.  With >= 3 packed stages per group the MMA issuer of block i does it
        // (it observes wready(i), which this group only signals after every thread has read the stage: no barrier or copy
        // issue inside this chain; 7.7 -> 7.5 us at M = 16).  With 1-2 stages per group that is too late — the group would
        // wait a full HBM round trip per block (M = 128: 10.7 -> 12.3 us measured) — so the group's leader refills right
        // after the whole group has read.
        if (!ISSUER_REFILL) {
          asm volatile("bar.sync %0, %1;" ::"r"(1 + gq), "r"(TG) : "memory");
          if (tl == 0 && i + PST < NI) load_weights(i + PST, s);
        }
        if (i >= WST) mbar_wait(bar_wempty + 8 * ws, ((i / WST) & 1) ^ 1);  // the MMA of block i - WST has read the stage
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int q = tl + TG * u;
          const int f_lo = (q >> 5) * 16 + g;
          uint4 lo[2], hi[2];
          Dequant<T, 4>::run(pv[u], s_lo[u], zl[u], s_hi[u], zh[u], lo, hi);
          const uint32_t rlo = sW + ws * C::W_BYTES + f_lo * 128;
          const uint32_t rhi = rlo + 8 * 128;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint32_t off = (((uint32_t)(2 * tt + c)) ^ (uint32_t)g) << 4;  // (row & 7) == g for rows f and f+8
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(rlo + off), "r"(lo[c].x), "r"(lo[c].y),
                         "r"(lo[c].z), "r"(lo[c].w)
                         : "memory");
            asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(rhi + off), "r"(hi[c].x), "r"(hi[c].y),
                         "r"(hi[c].z), "r"(hi[c].w)
                         : "memory");
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(bar_wready + 8 * ws);
      }
    } else {
      // 8-bit: a group covers the 128 feature rows x two 32-k halves of a block: thread -> rows tl + TG*r, both halves,
      // two uint4 (16 k each) per (row, half)
      constexpr int R = (128 / TG) > 0 ? 128 / TG : 1;  // (TG = 256 only exists for the 4-bit debug variant)
      for (int i = gq; i < NI; i += DQG) {
        const int kb = kb0 + i, s = i % PST, ws = i % WST;
        SZRaw sz[R][2];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int n = n0 + tl + TG * r;
#pragma unroll
          for (int j = 0; j < 2; ++j) sz[r][j] = load_sz<T, BITS, ASYM>(scales, qzeros, (2 * kb + j) >> gshc, n < N ? n : 0, N);
        }
        mbar_wait(bar_pfull + 8 * s, (i / PST) & 1);
        uint4 pvs[R][2][2];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int fr = tl + TG * r;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const uint4* pj = reinterpret_cast<const uint4*>(smem + (sPs(s) - smem_base) + j * C::P_CHUNK_BYTES);
#pragma unroll
            for (int h = 0; h < 2; ++h) pvs[r][j][h] = pj[((fr >> 5) * 2 + h) * 32 + (fr & 31)];
          }
        }
        if (!ISSUER_REFILL) {  // see the 4-bit path
          asm volatile("bar.sync %0, %1;" ::"r"(1 + gq), "r"(TG) : "memory");
          if (tl == 0 && i + PST < NI) load_weights(i + PST, s);
        }
        if (i >= WST) mbar_wait(bar_wempty + 8 * ws, ((i / WST) & 1) ^ 1);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int fr = tl + TG * r;
          const int n = n0 + fr;
          const int nsafe = n < N ? n : 0;
          const uint32_t brow = sW + ws * C::W_BYTES + fr * 128;
          const uint32_t sw = (uint32_t)(fr & 7);
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            int z = ZSYM;
            if (ASYM) z = (int)((sz[r][j].zw >> (BITS * (nsafe % PF))) & ((1u << BITS) - 1));
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              uint4 o[2];
              Dequant<T, 8>::run(pvs[r][j][h], sz[r][j].s, z, o);
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                const uint32_t addr = brow + (((uint32_t)(j * 4 + h * 2 + c) ^ sw) << 4);
                asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(o[c].x), "r"(o[c].y), "r"(o[c].z),
                             "r"(o[c].w)
                             : "memory");
              }
            }
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(bar_wready + 8 * ws);
      }
    }

    // ---- this rank's fp32 accumulator D[feature][token] -> its own shared memory, transposed: part[token][feature]
    // (the stage buffers are idle once bar_tfull fired: every load was consumed by an MMA that has completed)
    mbar_wait(bar_tfull, 0);
    tc_fence_after();
    asm volatile("griddepcontrol.wait;" ::: "memory");  // (already satisfied) orders the global stores below
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;       // two warps share a quarter: they split the 16-token column chunks
    // column chunk c of 16 tokens; with two weight sets the chunks NTOK/16 .. 2*NTOK/16-1 are set 1 (TMEM columns and
    // partial-tile rows continue seamlessly: part[set * NTOK + token][feature])
    const int nacc = nkb >= C::NISS ? C::NISS : (nkb >= 2 ? 2 : 1);  // accumulators in use (one per active MMA issuer)
    for (int c = half; c < NSETS * NTOK / 16; c += 2) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(tbase + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16), r);
      tmem_ld_wait();
      for (int ai = 1; ai < nacc; ++ai) {  // fixed order: deterministic
        uint32_t r2[16];
        tmem_ld_32x32b_x16(tbase + ((uint32_t)(q * 32) << 16) + (uint32_t)(ai * C::ACC_COLS + c * 16), r2);
        tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 16; ++v) r[v] = __float_as_uint(__uint_as_float(r[v]) + __uint_as_float(r2[v]));
      }
#pragma unroll
      for (int v = 0; v < 16; ++v)
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sW + (uint32_t)(c * 16 + v) * (MM_BF * 4) + (uint32_t)(q * 32 + lane) * 4),
                     "r"(r[v])
                     : "memory");
    }
    tc_fence_before();
  }
  // all ranks' partials are in place (cluster barrier = CTA barrier + cross-CTA release / acquire)
  __syncwarp();
  cluster_sync_all();
  if (warp >= 2 && warp < 10) {
    // rank z reduces token rows z, z + nrank, ... over all ranks through distributed shared memory; a warp owns one
    // token row at a time: 32 lanes x 4 features = the 128 features of the tile = 256 contiguous output bytes
    const int t = threadIdx.x - 64;
    const int chunk = t & 31;
    const int nc = n0 + chunk * 4;
    if (nc < N) {
      for (int tok = (int)crank + (int)nrank * (t >> 5); tok < M; tok += (int)nrank * MM_DQ_WARPS) {
        const uint32_t local = sW + (uint32_t)tok * (MM_BF * 4) + (uint32_t)chunk * 16;
        float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
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
          if (MODE == 1) {
            asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "r"(ra + (uint32_t)NTOK * (MM_BF * 4))
                         : "memory");
            acc2[0] += v.x;
            acc2[1] += v.y;
            acc2[2] += v.z;
            acc2[3] += v.w;
          }
        }
        if (MODE == 0) {
          if (bias != nullptr) {
            // reference order: round the matmul to the output dtype, then add bias (torch.py:337-342)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = E::to_f(E::from_f(acc[i])) + E::to_f(bias[nc + i]);
          }
          *reinterpret_cast<uint2*>(out + (size_t)tok * N + nc) =
              make_uint2(E::pack2(acc[0], acc[1]), E::pack2(acc[2], acc[3]));
        } else if (MODE == 1) {
          // the per-expert module loop of the reference model rounds at every module boundary
          // (act_fn(w1(x)) * w3(x) with 16-bit tensors): g = T(x W1), a = T(silu(g)), u = T(x W3), h = T(a * u)
          float hv[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float gq = E::to_f(E::from_f(acc[i])), uq = E::to_f(E::from_f(acc2[i]));
            const float aq = E::to_f(E::from_f(gq / (1.f + __expf(-gq))));
            hv[i] = aq * uq;
          }
          *reinterpret_cast<uint2*>(out + (size_t)(row0 + tok) * N + nc) =
              make_uint2(E::pack2(hv[0], hv[1]), E::pack2(hv[2], hv[3]));
        } else {
          // y = T(h W2) like the module, then the routing weight; kept in fp32 in the row's (token, k) slot — the k slots
          // of a token are summed (and rounded ONCE) by moe_combine_kernel: no atomics, deterministic
          const int pair = G.sorted_pairs[row0 + tok];
          const float w = G.pair_weights[pair];
          float4 o;
          o.x = w * E::to_f(E::from_f(acc[0]));
          o.y = w * E::to_f(E::from_f(acc[1]));
          o.z = w * E::to_f(E::from_f(acc[2]));
          o.w = w * E::to_f(E::from_f(acc[3]));
          *reinterpret_cast<float4*>(G.ypair + (size_t)pair * N + nc) = o;
        }
      }
    }
  }
  __syncwarp();
  cluster_sync_all();  // keep every rank's shared memory alive until all peers have read it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tbase, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
int make_x_tmap_box(CUtensorMap* map, const void* x, int M, int K, int dtype, int box_rows);  // b2q_gemm.cu

// split-K ranks (cluster size) for one launch: fill ~148 SMs, keep >= 4 k-blocks per rank, portable cluster size <= 8
int midm_ranks(int K, int N) {
  const int tiles = (N + MM_BF - 1) / MM_BF, nkb = K / MM_BK;
  int ks = 1;
  while (ks < 8 && tiles * ks * 2 <= 148 && nkb / (ks * 2) >= 4) ks *= 2;
  return ks;
}

constexpr int MM_DQG = 4;  // dequant groups = k-blocks dequantised concurrently

template <typename T, int BITS, bool ASYM, int NTOK, int PST, int WST, int MODE = 0, int DQG = MM_DQG>
static int launch_midm_t(const MmArgs& a, const void* x, int ks, const MoeArgs& G = MoeArgs{}, int x_rows = 0,
                         int grid_z = 1) {
  using C = MidCfg<BITS, NTOK, PST, WST, MODE>;
  if constexpr (BITS == 4 && MODE == 0 && DQG == MM_DQG) {
    if (env().midm_dqg1)  // debugging: all dequant warps on the same k-block
      return launch_midm_t<T, BITS, ASYM, NTOK, PST, WST, 0, 1>(a, x, ks, G, x_rows, grid_z);
  }
  CUtensorMap tmap;
  if (make_x_tmap_box(&tmap, x, MODE == 0 ? a.M : x_rows, a.K, a.dtype, NTOK) != 0) return -1;
  auto kern = midm_kernel<T, BITS, ASYM, NTOK, PST, WST, MODE, DQG>;
  static uint32_t smem_ok = 0;  // per-device bit mask (a process may drive several GPUs)
  if (int e = ensure_dyn_smem(kern, C::SMEM_BYTES, smem_ok, "b2q_midm")) return e;
  const int nkb = a.K / MM_BK;
  const int kpc = (nkb + ks - 1) / ks;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((a.N + MM_BF - 1) / MM_BF, ks, grid_z);
  cfg.blockDim = dim3(MM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
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
  return (int)cudaLaunchKernelEx(&cfg, kern, tmap, (const uint4*)a.packed, (const T*)a.scales,
                                 (const uint32_t*)a.qzeros, (const T*)a.bias, (T*)a.out, a.M, a.K, a.N, gemm_gshc(a),
                                 kpc, G);
}

bool midm_supported(const MmArgs& a) { return a.M >= 1 && a.M <= 128 && a.K % MM_BK == 0 && a.N % 32 == 0; }

// x: activations with act-order already applied (launch_gemm permutes into the workspace first)
int launch_midm(const MmArgs& a, const void* x) {
  int ks = env().midm_ks > 0 ? env().midm_ks : midm_ranks(a.K, a.N);
  if (ks > 8) ks = 8;
  const int nkb = a.K / MM_BK;
  while (ks > 1 && (ks - 1) * ((nkb + ks - 1) / ks) >= nkb) ks >>= 1;  // every rank needs at least one k-block
  const bool asym = a.qzeros != nullptr;
  // ring depths: BOTH must be multiples of the number of dequant groups, so that every use of a stage is served by the
  // SAME group.  mbarrier waits only distinguish the parity of a phase: with 6-stage rings and 4 groups (first version)
  // consecutive uses of a stage belonged to different groups, a group that ran two uses ahead saw the phase of use u - 2 as
  // "its" completed phase, consumed a stage that had not been refilled and corrupted the arrival counts — every
  // single-launch parity test and all three sanitizer tools passed, back-to-back launches at full size faulted
  // (profiles/r02_midm_notes.md).  8 dequantised stages = two per group; 4-8 packed stages (refilled by the group's own
  // leader); 4-8 activation stages in their own ring.
#define B2Q_MM_NTOK(T, BITS, AS)                                                                  \
  (a.M <= 16   ? launch_midm_t<T, BITS, AS, 16, (BITS == 4 ? 12 : 8), 8>(a, x, ks)                \
   : a.M <= 32 ? launch_midm_t<T, BITS, AS, 32, (BITS == 4 ? 12 : 8), 8>(a, x, ks)                \
   : a.M <= 64 ? launch_midm_t<T, BITS, AS, 64, (BITS == 4 ? 8 : 4), 8>(a, x, ks)                 \
               : launch_midm_t<T, BITS, AS, 128, 4, 8>(a, x, ks))
#define B2Q_MM_CASE(T)                                                          \
  (a.bits == 4 ? (asym ? B2Q_MM_NTOK(T, 4, true) : B2Q_MM_NTOK(T, 4, false))   \
               : (asym ? B2Q_MM_NTOK(T, 8, true) : B2Q_MM_NTOK(T, 8, false)))
  return a.dtype == 0 ? B2Q_MM_CASE(__half) : B2Q_MM_CASE(__nv_bfloat16);
#undef B2Q_MM_CASE
#undef B2Q_MM_NTOK
}

// ------------------------------------------------------------------------------------------------
// grouped launches for a MoE block (b2q_moe.cu): a = {x = expert-sorted activations [rows, K], packed / scales / qzeros =
// the STACKED tensors of all experts (expert stride = one expert's tensor), out = h [rows, N] (mode 1), M = token-box
// width hint (largest row count one expert is expected to get), N / K of ONE expert}
int midm_grouped_ranks(int K, int N, int active) {
  const int tiles = (N + MM_BF - 1) / MM_BF, nkb = K / MM_BK;
  int ks = 1;
  while (ks < 8 && tiles * active * ks * 2 <= 148 && nkb / (ks * 2) >= 4) ks *= 2;
  return ks;
}

int launch_midm_grouped(int mode, const MmArgs& a, const MoeGroupedArgs& g) {
  if (a.bits != 4 || a.K % MM_BK != 0 || a.N % 32 != 0 || (mode != 1 && mode != 2) || g.E < 1 || g.rows < 1) {
    set_error("b2q_moe: grouped launch needs bits=4, K %% 64 == 0, N %% 32 == 0 (bits=%d K=%d N=%d E=%d rows=%d)", a.bits,
              a.K, a.N, g.E, g.rows);
    return -1;
  }
  MoeArgs G = {};
  G.counts = g.counts;
  G.offsets = g.offsets;
  G.sorted_pairs = g.sorted_pairs;
  G.pair_weights = g.pair_weights;
  G.packed3 = (const uint4*)g.packed3;
  G.scales3 = g.scales3;
  G.qzeros3 = (const uint32_t*)g.qzeros3;
  G.ypair = g.ypair;
  // token-box width: every expert may receive up to `rows` rows; boxes of 16 serve decode-sized batches with one block
  // per expert, larger batches take 128-row blocks (CTAs of blocks beyond an expert's count exit immediately)
  const int ntok = g.rows <= 16 ? 16 : g.rows <= 32 ? 32 : g.rows <= 64 ? 64 : 128;
  G.tblocks = (g.rows + ntok - 1) / ntok;
  const int grid_z = g.E * G.tblocks;
  int ks = env().midm_ks > 0 ? env().midm_ks : midm_grouped_ranks(a.K, a.N, g.active > 0 ? g.active : 1);
  if (ks > 8) ks = 8;
  const int nkb = a.K / MM_BK;
  while (ks > 1 && (ks - 1) * ((nkb + ks - 1) / ks) >= nkb) ks >>= 1;
  const bool asym = a.qzeros != nullptr;
#define B2Q_MG_NTOK(T, AS, MODE)                                                             \
  (ntok == 16   ? launch_midm_t<T, 4, AS, 16, 12, 8, MODE>(a, a.x, ks, G, g.rows, grid_z)    \
   : ntok == 32 ? launch_midm_t<T, 4, AS, 32, 12, 8, MODE>(a, a.x, ks, G, g.rows, grid_z)    \
   : ntok == 64 ? launch_midm_t<T, 4, AS, 64, 8, 8, MODE>(a, a.x, ks, G, g.rows, grid_z)     \
                : launch_midm_t<T, 4, AS, 128, 4, 8, MODE>(a, a.x, ks, G, g.rows, grid_z))
#define B2Q_MG_CASE(T)                                                                       \
  (mode == 1 ? (asym ? B2Q_MG_NTOK(T, true, 1) : B2Q_MG_NTOK(T, false, 1))                   \
             : (asym ? B2Q_MG_NTOK(T, true, 2) : B2Q_MG_NTOK(T, false, 2)))
  return a.dtype == 0 ? B2Q_MG_CASE(__half) : B2Q_MG_CASE(__nv_bfloat16);
#undef B2Q_MG_CASE
#undef B2Q_MG_NTOK
}

}  // namespace b2q
