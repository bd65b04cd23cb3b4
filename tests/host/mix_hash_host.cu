// Host-side known-answer generator for tests/test_data_path_cpu.py: prints the product's counter-based mixes
// (fer_vit_b200/csrc/common.cuh: mix_hash64 / mix_hash / drop_threshold are __host__ __device__) for the (seed, site,
// index) triples given on the command line, so the oracle's numpy restatement is pinned on the C source without a GPU.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "common.cuh"

int main(int argc, char** argv) {
  for (int i = 1; i + 2 < argc; i += 3) {
    const unsigned long long seed = strtoull(argv[i], nullptr, 0);
    const unsigned int site = (unsigned int)strtoul(argv[i + 1], nullptr, 0);
    const unsigned long long idx = strtoull(argv[i + 2], nullptr, 0);
    printf("%llu %u\n", (unsigned long long)fervit::mix_hash64(seed, site, idx), fervit::mix_hash(seed, site, idx));
  }
  {  // the dropout generator at and beyond 32-bit indices
    const unsigned long long idxs[4] = {0ull, 77ull, 0xFFFFFFFFull, (1ull << 40) + 7ull};
    printf("H");
    for (int k = 0; k < 4; ++k) printf(" %u", fervit::drop_hash(0xC0FFEEull, 0x4C41u + k, idxs[k]));
    printf("\n");
  }
  {  // keep-mask of 64 consecutive elements of one dropout site at p = 0.1 and p = 0.5, as bit strings
    const float ps[2] = {0.1f, 0.5f};
    for (int k = 0; k < 2; ++k) {
      printf("K ");
      for (unsigned long long i = 0; i < 64; ++i)
        putchar(fervit::drop_keep(12345ull, 9u, 1000ull + i, fervit::drop_threshold(ps[k])) ? '1' : '0');
      printf("\n");
    }
  }
  {  // worst absolute error of the MUFU-free GELU polynomials of the tensor-core epilogues against the erf forms
    double e_in = 0, e_tail = 0, d_in = 0, d_tail = 0;
    for (int i = -80000; i <= 80000; ++i) {
      const double u = i * 1e-4;
      const double cdf = 0.5 * (1.0 + erf(u / sqrt(2.0)));
      const double g = u * cdf, d = cdf + u * exp(-0.5 * u * u) / sqrt(2.0 * 3.14159265358979323846);
      const double eg = fabs((double)fervit::gelu_fwd_poly((float)u) - g);
      const double ed = fabs((double)fervit::gelu_bwd_poly((float)u) - d);
      if (fabs(u) <= 4.0) { if (eg > e_in) e_in = eg; if (ed > d_in) d_in = ed; }
      else { if (eg > e_tail) e_tail = eg; if (ed > d_tail) d_tail = ed; }
    }
    printf("G %.3e %.3e %.3e %.3e\n", e_in, e_tail, d_in, d_tail);
  }
  {  // the packed (FFMA2) forward epilogue: one min on u^2 + saturating cdf, value and derivative in one pass
    double e_in = 0, e_tail = 0, d_in = 0, d_tail = 0;
    for (int i = -80000; i <= 80000; ++i) {
      const double u = i * 1e-4;
      const double cdf = 0.5 * (1.0 + erf(u / sqrt(2.0)));
      const double g = u * cdf, d = cdf + u * exp(-0.5 * u * u) / sqrt(2.0 * 3.14159265358979323846);
      float gg, dd;
      fervit::gelu_fwd_deriv_packed_mirror((float)u, gg, dd);
      const double eg = fabs((double)gg - g), ed = fabs((double)dd - d);
      if (fabs(u) <= 4.0) { if (eg > e_in) e_in = eg; if (ed > d_in) d_in = ed; }
      else { if (eg > e_tail) e_tail = eg; if (ed > d_tail) d_tail = ed; }
    }
    printf("E %.3e %.3e %.3e %.3e\n", e_in, e_tail, d_in, d_tail);
  }
  printf("T %u %u %u\n", fervit::drop_threshold(0.1f), fervit::drop_threshold(0.5f), fervit::drop_threshold(0.999f));
  return 0;
}
