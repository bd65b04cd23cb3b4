// Host-side check of the stream-K schedule of the CTA-pair GEMM (fer_vit_b200/csrc/gemm_tc2_sched.cuh, the very code
// the kernel runs): for every (tile count, k-blocks, pair count) it replays what each pair would do and verifies the
// invariants the device protocol relies on. Prints "OK <cases>" or the first violation.
#include <cstdio>
#include <cstring>
#include <vector>
#include "gemm_tc2_sched.cuh"

using namespace fervit::tc2;

static int fail(const char* what, int units, int kb, int mp, int pair) {
  printf("FAIL %s units=%d kb=%d max_pairs=%d pair=%d\n", what, units, kb, mp, pair);
  return 1;
}

int main() {
  constexpr int BN = 256;
  int cases = 0;
  const int pair_counts[3] = {74, 66, 8};
  const int kbs[6] = {16, 17, 24, 36, 48, 64};
  for (int mpi = 0; mpi < 3; ++mpi)
    for (int ki = 0; ki < 6; ++ki)
      for (int units = 1; units <= 260; ++units) {
        const int mp = pair_counts[mpi], kb = kbs[ki];
        Params p;
        memset(&p, 0, sizeof(p));
        p.n_blocks = 3;
        p.pair_m_blocks = (units + 2) / 3;
        p.full_units = units / mp * mp;
        const int rest = units - p.full_units;
        p.split = 1;
        p.virt_units = p.full_units;
        if (rest == 0) continue;
        const int q = (int)(((long long)rest * kb + mp - 1) / mp);   // same rule as tc2::launch
        if (q + 4 > kb) continue;
        p.sk_q = q;
        p.sk_tiles = rest;
        int pairs = p.full_units < mp ? p.full_units : mp;
        const int sk_pairs = (int)(((long long)rest * kb + q - 1) / q);
        if (sk_pairs > pairs) pairs = sk_pairs;
        if (pairs > mp) return fail("more pairs than the GPU has", units, kb, mp, -1);
        std::vector<int> unit_seen(p.full_units, 0), cover((size_t)rest * kb, 0), partials(rest, 0), first_w(rest, 1 << 30),
            last_w(rest, -1), owner(rest, -1), owner_ka(rest, 0);
        for (int pair = 0; pair < pairs; ++pair) {
          Item w;
          int nseg = 0, npart = 0, prev_tile = -1;
          bool in_sk = false;
          for (int it = 0; get_item<BN, true>(p, pair, pairs, kb, it, w); ++it) {
            if (w.sk_tile < 0) {
              if (in_sk) return fail("ordinary unit after a stream-K segment", units, kb, mp, pair);
              const int u = pair + it * pairs;
              if (u >= p.full_units || w.ka != 0 || w.ke != kb) return fail("bad ordinary unit", units, kb, mp, pair);
              unit_seen[u]++;
              continue;
            }
            in_sk = true;
            ++nseg;
            if (w.sk_tile >= rest || w.ka < 0 || w.ke > kb || w.ka >= w.ke) return fail("bad segment", units, kb, mp, pair);
            if (w.t.n_blk != (p.full_units + w.sk_tile) % p.n_blocks || w.t.pm != (p.full_units + w.sk_tile) / p.n_blocks)
              return fail("tile decode", units, kb, mp, pair);
            for (int k = w.ka; k < w.ke; ++k) cover[(size_t)w.sk_tile * kb + k]++;
            if (w.ke < kb) {   // partial writer
              ++npart;
              partials[w.sk_tile]++;
              if (pair < first_w[w.sk_tile]) first_w[w.sk_tile] = pair;
              if (pair > last_w[w.sk_tile]) last_w[w.sk_tile] = pair;
              if (nseg == 2) return fail("a head must be computed before the tail", units, kb, mp, pair);
            } else {
              owner[w.sk_tile] = pair;
              owner_ka[w.sk_tile] = w.ka;
              if (nseg == 2 && (prev_tile != w.sk_tile + 1)) return fail("tail is not the tile before the head", units, kb, mp, pair);
            }
            prev_tile = w.sk_tile;
          }
          if (nseg > 2) return fail("more than two segments", units, kb, mp, pair);
          if (npart > 1) return fail("a pair writes more than one partial (one scratch slot per pair)", units, kb, mp, pair);
        }
        for (int u = 0; u < p.full_units; ++u)
          if (unit_seen[u] != 1) return fail("ordinary unit not covered exactly once", units, kb, mp, u);
        for (size_t i = 0; i < cover.size(); ++i)
          if (cover[i] != 1) return fail("k-block not covered exactly once", units, kb, mp, (int)(i / kb));
        for (int t = 0; t < rest; ++t) {
          if (owner[t] < 0) return fail("tile without an owner", units, kb, mp, t);
          const int sk_first = (int)(((long long)t * kb) / q), sk_n = owner[t] - sk_first;   // what the owner computes
          if (owner_ka[t] == 0) {
            if (partials[t] != 0) return fail("full tile with partials", units, kb, mp, t);
            continue;
          }
          if (sk_n != partials[t]) return fail("owner expects a different number of partials", units, kb, mp, t);
          if (first_w[t] != sk_first || last_w[t] != owner[t] - 1) return fail("partial writers are not sk_first..owner-1", units, kb, mp, t);
        }
        ++cases;
      }
  // ---- tail splitting without stream-K: every column of every tile exactly once, slices only in the last round ----
  int split_cases = 0;
  for (int mpi = 0; mpi < 3; ++mpi)
    for (int units = 1; units <= 400; ++units) {
      const int mp = pair_counts[mpi], kb = 12;
      Params p;
      memset(&p, 0, sizeof(p));
      p.n_blocks = 9;
      p.pair_m_blocks = (units + 8) / 9;
      p.full_units = units / mp * mp;
      const int rest = units - p.full_units;
      p.split = 1;
      if (p.full_units > 0 && rest > 0) {            // same rule as tc2::launch (bf16 kinds: slices of >= 64 columns)
        if (rest * 4 <= mp) p.split = 4;
        else if (rest * 2 <= mp) p.split = 2;
      }
      p.virt_units = p.full_units + rest * p.split;
      const int pairs = p.virt_units < mp ? p.virt_units : mp;
      std::vector<int> col((size_t)units * BN, 0);
      for (int pair = 0; pair < pairs; ++pair) {
        Item w;
        for (int it = 0; get_item<BN, false>(p, pair, pairs, kb, it, w); ++it) {
          const int tile = w.t.pm * p.n_blocks + w.t.n_blk;
          if (tile >= units || w.ka != 0 || w.ke != kb || w.sk_tile != -1) return fail("bad unit", units, kb, mp, pair);
          if (w.t.width != BN && tile < p.full_units) return fail("slice outside the last round", units, kb, mp, pair);
          if (w.t.width * p.split != BN && w.t.width != BN) return fail("slice width", units, kb, mp, pair);
          for (int c = w.t.col_off; c < w.t.col_off + w.t.width; ++c) col[(size_t)tile * BN + c]++;
        }
      }
      for (size_t i = 0; i < col.size(); ++i)
        if (col[i] != 1) return fail("column not covered exactly once", units, kb, mp, (int)(i / BN));
      ++split_cases;
    }
  printf("OK %d %d\n", cases, split_cases);
  return 0;
}
