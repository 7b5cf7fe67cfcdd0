// b2q_streamk.h — work decomposition of the stream-K prefill kernel (b2q_gemm2s.cu), shared by host and device so
// that the host-side tests (b2q_debug_gemm_plan) exercise exactly the code the kernel runs.
//
// The output is cut into 256 x 256 tiles (tile = tm + TM * tn, token tiles fastest); every tile needs `nkb` k-blocks
// of 64.  P CTA pairs are resident.  Plain round-robin ("data parallel") leaves the last wave partly empty: 128 tiles
// on 74 pairs run as 2 waves at 86 % (profiles/r01_gemm_notes.md).  Here the LAST sk_tiles tiles are treated as one
// stream of sk_tiles * nkb k-block units cut into P equal contiguous segments (one per pair, processed FIRST), the
// other tiles stay data parallel (whole tiles, round-robin).  A segment that starts inside a tile produces a partial
// accumulator: the pair that holds the tile's FIRST k-block owns the tile and adds the partials of the pairs that
// follow it (they reach that piece at the START of their segment, so the partial is normally ready long before the
// owner asks for it).
#pragma once

#if defined(__CUDACC__)
#define B2Q_HD __host__ __device__ __forceinline__
#else
#define B2Q_HD inline
#endif

namespace b2q {

struct SkPlan {
  int tiles;     // TM * TN
  int P;         // CTA pairs of the launch
  int nkb;       // k-blocks (64 k) per tile
  int dp_tiles;  // tiles [0, dp_tiles): data parallel, pair p owns p, p + P, ...
  int sk_tiles;  // tiles [dp_tiles, tiles): stream-K
};

struct SkItem {
  int tile, kb0, kb1;  // k-blocks [kb0, kb1) of `tile`
};

B2Q_HD long long sk_units(const SkPlan& pl) { return (long long)pl.sk_tiles * pl.nkb; }
B2Q_HD long long sk_begin(const SkPlan& pl, int p) { return sk_units(pl) * p / pl.P; }

// Host: choose P and the stream-K region.  max_pairs = SM pairs available (74 on B200).
B2Q_HD SkPlan sk_make_plan(int tiles, int nkb, int max_pairs) {
  SkPlan pl;
  pl.tiles = tiles;
  pl.nkb = nkb;
  int P = max_pairs;
  if (tiles < P) {
    // fewer tiles than pairs: split every tile, but keep at least a quarter tile (and 4 k-blocks) per pair
    long long cap = (long long)tiles * 4;
    const long long cap2 = (long long)tiles * nkb / 4;
    if (cap2 < cap) cap = cap2;
    if (cap < tiles) cap = tiles;
    if (cap < P) P = (int)cap;
  }
  pl.P = P;
  const int rem = tiles % P;
  if (rem == 0) {
    pl.sk_tiles = 0;  // whole waves: nothing to balance
  } else if (tiles < P) {
    pl.sk_tiles = tiles;
  } else {
    pl.sk_tiles = rem + P;  // "two-tile" stream-K: every pair gets between one and two tiles' worth of k-blocks
    if (pl.sk_tiles > tiles) pl.sk_tiles = tiles;
  }
  pl.dp_tiles = tiles - pl.sk_tiles;
  return pl;
}

// The items of pair p, in processing order: its stream-K segment first, then its data-parallel tiles.
struct SkIter {
  SkPlan pl;
  long long u, uend;
  int dp;
  B2Q_HD SkIter(const SkPlan& plan, int p) : pl(plan) {
    u = sk_begin(plan, p);
    uend = sk_begin(plan, p + 1);
    dp = p;
  }
  B2Q_HD bool next(SkItem& it) {
    if (u < uend) {
      const int ts = (int)(u / pl.nkb);
      const int kb0 = (int)(u - (long long)ts * pl.nkb);
      long long e = (long long)(ts + 1) * pl.nkb;
      if (e > uend) e = uend;
      it.tile = pl.dp_tiles + ts;
      it.kb0 = kb0;
      it.kb1 = kb0 + (int)(e - u);
      u = e;
      return true;
    }
    if (dp < pl.dp_tiles) {
      it.tile = dp;
      it.kb0 = 0;
      it.kb1 = pl.nkb;
      dp += pl.P;
      return true;
    }
    return false;
  }
};

// For the owner of a split tile (item with kb0 == 0 and kb1 < nkb of pair p): the pairs q > p that hold the rest of the
// tile are exactly first_q .. first_q + count - 1 (pairs with an empty segment in between are skipped by the caller
// through sk_begin(q) == sk_begin(q + 1)).  Returns the LAST pair that touches the tile.
B2Q_HD int sk_last_contributor(const SkPlan& pl, int tile) {
  const long long last_unit = (long long)(tile - pl.dp_tiles + 1) * pl.nkb - 1;
  // smallest q with sk_begin(q + 1) > last_unit
  int q = (int)((last_unit * pl.P) / sk_units(pl));
  while (q > 0 && sk_begin(pl, q) > last_unit) --q;
  while (sk_begin(pl, q + 1) <= last_unit) ++q;
  return q;
}

}  // namespace b2q
