// Work decomposition of the CTA-pair GEMM (gemm_tc2.cu): kernel parameters, tile decoding (n-fastest order, tail
// splitting) and the per-pair work items including the stream-K segments. Host-callable so that the schedule's
// invariants (every k-block of every tile covered exactly once, heads before tails, one partial per pair, the
// owner's expected partial count) are checked on the CPU by tests/host/streamk_sched_host.cu.
#pragma once
#include <cstddef>
#include "common.cuh"   // Dropout

namespace fervit {
namespace tc2 {

struct Params {
  int M, N, K;
  int pair_m_blocks, n_blocks;
  const float* bias;
  const float* alpha_ptr;
  float alpha;
  int has_out, has_z, has_f32, has_res;
  int pre_is_deriv;  // forward activations: the Z tile receives act'(pre) instead of pre
  // Tail splitting: the tiles of the last, partial round (units - full_units of them) are cut into `split` column
  // slices of BN / split columns each, so that round costs 1 / split of a tile time (epilogue included) instead of a
  // whole one: at batch 256 the QKV GEMM has 2.31 rounds of tiles, fc1 3.08.
  int full_units, split, virt_units;
  // Stream-K over the last, partial round (plain epilogues, K >= 1024): its `sk_tiles` tiles (numbered from
  // full_units) are laid end to end as sk_tiles * (K / 64) k-blocks and every pair takes `sk_q` consecutive ones, so
  // the round costs sk_q k-blocks instead of K / 64 (57 tiles on 74 pairs: 37 instead of 48). A pair's range covers the
  // tail of one tile and/or the head of the next; it computes the head FIRST, dumps that fp32 partial accumulator to
  // `sk_ws` (slot = pair) and bumps the tile's flag; the pair holding the tile's last k-blocks adds the partials of
  // the pairs before it in ascending order (deterministic) and runs the normal epilogue. No pair ever waits for a
  // pair that can wait itself, so the grid cannot deadlock. 0 = off.
  // Status: correct and deterministic (tests), but at batch 256 it still LOSES: M=4864 N=768 K=3072 takes 23.9 us
  // plain and 26.8 us with stream-K (K=2304: 19.2 / 23.1; tools/streamk_bench.py). It saves 11 of 48 k-blocks (~3 us of
  // MMA time) and pays ~6 us for dumping and re-reading 128 KB of partials per CTA, the re-read sitting on the
  // epilogue's critical path. (A row-major scratch layout cost 21 us: 16-byte accesses at a 1 KB stride; the
  // lane-interleaved one is coalesced.) It is therefore opt-in - a caller-provided scratch, or FERVIT_GEMM_STREAMK=1
  // for the plans - until the owner prefetches the partial chunks into shared-memory staging (cp.async) before it
  // waits for its accumulator, which takes the re-read off the critical path ("next" 1). It also gives up
  // batch invariance of the results (the summation order then depends on the tile count).
  int sk_q, sk_tiles;
  float* sk_ws;   // [pairs][2][BM][BN] fp32
  int* sk_flags;  // [sk_tiles][2] arrival counters, zero between launches
  int debug;  // timing experiments only (results are garbage): 1 no TMA loads, 2 no MMAs, 4 no epilogue, 8 record clocks
  // in-kernel launch timer (fervit_gemm_prof): {min over CTAs of %globaltimer once the grid dependency has resolved,
  // max over CTAs at exit}; null = off. Works inside CUDA-graph replays, where host-side events cannot sit.
  unsigned long long* prof;
  // folded LayerNorm (common.cuh: Epilogue): consumer and producer side
  const float* ln_part; const float* ln_mref; const float* ln_cs; float* ln_mean; float* ln_rstd;
  float ln_eps; int ln_parts;
  float* lnp_part; const float* lnp_mref;
  // counter-based dropout on the epilogue's value, before the residual add (nn.TransformerEncoderLayer's dropout1 /
  // dropout / dropout2 of the post-norm models); threshold 0 = off. Kernels are instantiated with and without it.
  Dropout drop;
};



struct TileRef { int pm, n_blk, col_off, width; };
template <int BN>
__host__ __device__ __forceinline__ TileRef decode_unit(const Params& p, int v) {
  TileRef t;
  int tile = v;
  t.col_off = 0;
  t.width = BN;
  if (v >= p.full_units) {
    const int k = (v - p.full_units) / p.split, q = (v - p.full_units) % p.split;
    tile = p.full_units + k;
    t.width = BN / p.split;
    t.col_off = q * t.width;
  }
  // n fastest: the pairs running at one time share a band of A rows across all n-blocks, so A (the big operand:
  // activations) streams from HBM once and the weights (<= 14 MB) stay in L2. With m fastest, A was re-read once per
  // n-block as soon as it outgrew the 126 MB L2 (K = 3072 at batch >= 1024: TMA+MMA 143 us against 107 us of MMAs
  // alone, profiles/).
  t.n_blk = tile % p.n_blocks;
  t.pm = tile / p.n_blocks;
  return t;
}

// Work item `it` of a pair: first its round-robin units over all of K, then (stream-K) its one or two k-block segments
// of the last round's tiles. Returns false when the pair is done. sk_tile = -1 for ordinary units.
struct Item { TileRef t; int ka, ke, sk_tile; };
template <int BN, bool SK>
__host__ __device__ __forceinline__ bool get_item(const Params& p, int pair_id, int num_pairs, int total_kb, int it, Item& w) {
  const int u = pair_id + it * num_pairs;
  if (u < p.virt_units) {
    w.t = decode_unit<BN>(p, u);
    w.ka = 0;
    w.ke = total_kb;
    w.sk_tile = -1;
    return true;
  }
  if (!SK) return false;   // the ordinary instantiation is exactly the round-robin loop over units
  const int mine = pair_id < p.virt_units ? (p.virt_units - pair_id + num_pairs - 1) / num_pairs : 0;  // units before
  const int si = it - mine;
  const long long lo = (long long)pair_id * p.sk_q, tot = (long long)p.sk_tiles * total_kb;
  long long hi = lo + p.sk_q;
  if (hi > tot) hi = tot;
  if (lo >= hi) return false;
  const int t1 = (int)(lo / total_kb), a1 = (int)(lo % total_kb);
  int e1 = a1 + (int)(hi - lo);
  if (e1 > total_kb) e1 = total_kb;
  const int rem = (int)(hi - lo) - (e1 - a1);     // k-blocks that spill into the next tile: its head, done first
  int tile, ka, ke;
  if (rem > 0) {
    if (si == 0) { tile = t1 + 1; ka = 0; ke = rem; }
    else if (si == 1) { tile = t1; ka = a1; ke = e1; }
    else return false;
  } else {
    if (si != 0) return false;
    tile = t1; ka = a1; ke = e1;
  }
  w.sk_tile = tile;
  w.ka = ka;
  w.ke = ke;
  w.t.col_off = 0;
  w.t.width = BN;
  w.t.n_blk = (p.full_units + tile) % p.n_blocks;
  w.t.pm = (p.full_units + tile) / p.n_blocks;
  return true;
}

}  // namespace tc2
}  // namespace fervit
