// Host-side known-answer generator for tests/test_data_path_cpu.py: prints the product's counter-based mixes
// (fer_vit_b200/csrc/common.cuh: mix_hash64 / mix_hash / drop_threshold are __host__ __device__) for the (seed, site,
// index) triples given on the command line, so the oracle's numpy restatement is pinned on the C source without a GPU.
#include <cstdio>
#include <cstdlib>
#include "common.cuh"

int main(int argc, char** argv) {
  for (int i = 1; i + 2 < argc; i += 3) {
    const unsigned long long seed = strtoull(argv[i], nullptr, 0);
    const unsigned int site = (unsigned int)strtoul(argv[i + 1], nullptr, 0);
    const unsigned long long idx = strtoull(argv[i + 2], nullptr, 0);
    printf("%llu %u\n", (unsigned long long)fervit::mix_hash64(seed, site, idx), fervit::mix_hash(seed, site, idx));
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
  printf("T %u %u %u\n", fervit::drop_threshold(0.1f), fervit::drop_threshold(0.5f), fervit::drop_threshold(0.999f));
  return 0;
}
