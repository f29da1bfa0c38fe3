/* The product's pattern header (csrc/gsb_pattern.h, the one the sm_100a kernels include) compiled for the host:
 * prints mix32 values and pattern words so tests/test_oracle.py can hold it against the numpy and C oracles and the
 * golden fixture without a GPU.   usage: pattern_host mix <u32>... | word <first_word> <n> <seed> */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gsb_pattern.h"

int main(int argc, char **argv) {
  int i;
  if (argc >= 2 && strcmp(argv[1], "mix") == 0) {
    for (i = 2; i < argc; i++) printf("%" PRIu32 "\n", gsb_mix32((uint32_t)strtoull(argv[i], NULL, 10)));
    return 0;
  }
  if (argc == 5 && strcmp(argv[1], "word") == 0) {
    const uint64_t first = strtoull(argv[2], NULL, 10), n = strtoull(argv[3], NULL, 10);
    const uint32_t key = gsb_seed_key((uint32_t)strtoull(argv[4], NULL, 10));
    uint64_t w;
    for (w = first; w < first + n; w++) {
      uint32_t l[4];
      gsb_pattern_word(w, key, l);
      printf("%" PRIu32 " %" PRIu32 " %" PRIu32 " %" PRIu32 "\n", l[0], l[1], l[2], l[3]);
    }
    return 0;
  }
  return 64;
}
