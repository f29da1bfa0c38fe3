/*
 * gsb_pattern.h — the probe pattern, shared by the sm_100a kernels and the host side of the shim.
 *
 * The reference has no kernel ("CUDA kernels in the reference: NONE"), so the probe's
 * data pattern is defined here and restated independently in oracle/probe_oracle.c and
 * oracle/probe_oracle.py; the three must agree bit for bit (tests/test_probe_gpu.py).
 *
 * Unit = one 16-byte word (four 32-bit lanes), addressed by its ABSOLUTE index w in the arena
 * (byte offset / 16), so a window probe and a full walk see the same expected bytes:
 *
 *   key(w)        = mix32( lo32(w) ^ hi32(w) * 0x9E3779B1 )           -- address-unique ("own address")
 *   lane_l(w, s)  = ( key(w) * K[l] + C[l] ) ^ mix32( s ^ 0xA5A5A5A5 )  l = 0..3
 *
 * mix32 is the murmur3 finaliser (a bijection on 32 bits): two seeds never give the same word, so a
 * word that did not take a refill is always a mismatch; the address term catches aliasing. Integer
 * only (~20 ALU ops per 16 B), far below the 175 lane-instructions per word the SMs have at the HBM
 * roofline (DESIGN.md §kernel budget).
 */
#ifndef GSB_PATTERN_H_
#define GSB_PATTERN_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define GSB_HD __host__ __device__ __forceinline__
#else
#define GSB_HD static inline
#endif

#define GSB_PAT_K0 0x9E3779B1u
#define GSB_PAT_K1 0x85EBCA77u
#define GSB_PAT_K2 0xC2B2AE3Du
#define GSB_PAT_K3 0x27D4EB2Fu
#define GSB_PAT_C0 0x165667B1u
#define GSB_PAT_C1 0xD3A2646Cu
#define GSB_PAT_C2 0xFD7046C5u
#define GSB_PAT_C3 0xB55A4F09u
#define GSB_PAT_SEED_SALT 0xA5A5A5A5u

GSB_HD uint32_t gsb_mix32(uint32_t h) {
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}

GSB_HD uint32_t gsb_word_key(uint64_t w) {
  return gsb_mix32((uint32_t)w ^ ((uint32_t)(w >> 32) * GSB_PAT_K0));
}

GSB_HD uint32_t gsb_seed_key(uint32_t seed) { return gsb_mix32(seed ^ GSB_PAT_SEED_SALT); }

/* lanes[0..3] of word w under an already-mixed seed key */
GSB_HD void gsb_pattern_word(uint64_t w, uint32_t seed_key, uint32_t lanes[4]) {
  const uint32_t m = gsb_word_key(w);
  lanes[0] = (m * GSB_PAT_K0 + GSB_PAT_C0) ^ seed_key;
  lanes[1] = (m * GSB_PAT_K1 + GSB_PAT_C1) ^ seed_key;
  lanes[2] = (m * GSB_PAT_K2 + GSB_PAT_C2) ^ seed_key;
  lanes[3] = (m * GSB_PAT_K3 + GSB_PAT_C3) ^ seed_key;
}

#endif /* GSB_PATTERN_H_ */
