/*
 * probe_oracle.c — C restatement of the probe spec (DESIGN.md "probe pattern"); TEST INFRASTRUCTURE
 * ONLY (tests/, smoke(), bench.py's cpu_baseline). Independent of csrc/gsb_pattern.h on purpose:
 * constants and arithmetic are written out again from the spec, not included.
 *
 * The reference has nothing to restate here (it never touches HBM: pkg/gpu/nvidia/nvidia.go:100-152
 * is an NVML event wait). This exists so the numpy oracle (probe_oracle.py) has a second,
 * differently-written witness, and so larger windows can be checked in seconds (plain C; the omp pragmas are inert unless built with -fopenmp).
 */
#include <stdint.h>
#include <stddef.h>

static inline uint32_t mix32(uint32_t h) {
  h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
  return h;
}
static const uint32_t K[4] = {0x9E3779B1u, 0x85EBCA77u, 0xC2B2AE3Du, 0x27D4EB2Fu};
static const uint32_t C[4] = {0x165667B1u, 0xD3A2646Cu, 0xFD7046C5u, 0xB55A4F09u};

/* out[4*n_words] = pattern of words [first_word, first_word+n_words) under `seed` */
void po_pattern(uint64_t first_word, uint64_t n_words, uint32_t seed, uint32_t *out) {
  const uint32_t sk = mix32(seed ^ 0xA5A5A5A5u);
#pragma omp parallel for schedule(static)
  for (uint64_t i = 0; i < n_words; i++) {
    const uint64_t w = first_word + i;
    const uint32_t m = mix32((uint32_t)w ^ ((uint32_t)(w >> 32) * 0x9E3779B1u));
    for (int l = 0; l < 4; l++) out[4 * i + l] = (m * K[l] + C[l]) ^ sk;
  }
}

/* checksums of the pattern itself, without materialising it: res = {xor, sum} */
void po_pattern_checksums(uint64_t first_word, uint64_t n_words, uint32_t seed, uint32_t res[2]) {
  const uint32_t sk = mix32(seed ^ 0xA5A5A5A5u);
  uint32_t x = 0, s = 0;
#pragma omp parallel for schedule(static) reduction(^ : x) reduction(+ : s)
  for (uint64_t i = 0; i < n_words; i++) {
    const uint64_t w = first_word + i;
    const uint32_t m = mix32((uint32_t)w ^ ((uint32_t)(w >> 32) * 0x9E3779B1u));
    for (int l = 0; l < 4; l++) {
      const uint32_t v = (m * K[l] + C[l]) ^ sk;
      x ^= v;
      s += v;
    }
  }
  res[0] = x;
  res[1] = s;
}

/* verify observed bytes: res = {xor, sum, mismatch_words, mismatch_bits}; returns first bad word or ~0 */
uint64_t po_verify(const uint32_t *obs, uint64_t first_word, uint64_t n_words, uint32_t seed_expect,
                   uint64_t res[4]) {
  const uint32_t sk = mix32(seed_expect ^ 0xA5A5A5A5u);
  uint32_t x = 0, s = 0;
  uint64_t mw = 0, mb = 0, first = ~0ull;
#pragma omp parallel for schedule(static) reduction(^ : x) reduction(+ : s, mw, mb) reduction(min : first)
  for (uint64_t i = 0; i < n_words; i++) {
    const uint64_t w = first_word + i;
    const uint32_t m = mix32((uint32_t)w ^ ((uint32_t)(w >> 32) * 0x9E3779B1u));
    uint32_t any = 0;
    for (int l = 0; l < 4; l++) {
      const uint32_t v = obs[4 * i + l];
      const uint32_t d = v ^ ((m * K[l] + C[l]) ^ sk);
      x ^= v;
      s += v;
      any |= d;
      mb += (uint64_t)__builtin_popcount(d);
    }
    if (any) {
      mw++;
      if (w < first) first = w;
    }
  }
  res[0] = x; res[1] = s; res[2] = mw; res[3] = mb;
  return first;
}
