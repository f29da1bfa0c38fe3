/*
 * hbm_probe_sm100a.cu — the per-device HBM probe kernels (sm_100a only).
 *
 * Replaces the reference's passive health half — nvml.WaitForEvent in watchXIDs
 * (pkg/gpu/nvidia/nvidia.go:100-152), which never touches HBM — with an active walk of the arena:
 * every 16-byte word of the window is loaded once (compared with the pattern of the generation
 * that last wrote it) and stored once (the next generation's pattern). Algorithmic traffic of a
 * VERIFY_REFILL launch = 2 * window bytes; bound = HBM bandwidth. Integer only,
 * no tensor cores (this is not a contraction).
 *
 * Three data paths, same results bit for bit (tests/test_probe_gpu.py):
 *   DIRECT   ld.global.v4 -> registers -> st.global.v4          (control: no staging)
 *   CPASYNC  cp.async 16 B (LDGSTS) -> shared ring -> ld.shared.v4 -> st.global.v4
 *   BULK     cp.async.bulk (TMA 1-D, UBLKCP) + mbarrier -> shared ring -> ld.shared.v4 /
 *            st.shared.v4 -> cp.async.bulk shared->global
 * All are persistent grid-stride kernels over fixed-size tiles, grid = resident CTAs/SM x #SMs.
 *
 * What ships (default build): the two kernels GSB_VARIANT_AUTO ever launches — BULKD 32 KiB x 3 (FILL,
 * VERIFY_REFILL) and BULK 32 KiB x 3 (VERIFY) — plus ONE shape each of DIRECT, CPASYNC and BULKW, kept as
 * independent data paths for the parity tests (five implementations, one answer). The tile-shape / cache-operator /
 * L2-hint sweeps of round 1 (profiles/sweep_r01_*.json: 72 kernel instantiations, a 3.9 MB library) are compiled only
 * with -DGSB_LAB=1 (build.sh lab -> libgpushare_b200_lab.so, used by tools/sweep_r02.py); in the default build the
 * GSB_*_CFG / GSB_DIRECT_FLAVOR / GSB_L2_HINT / GSB_DYN_FILL knobs are ignored.
 *
 * Reduction: per-thread registers -> warp shuffles -> one gsb_partial slot per CTA (plain stores,
 * no atomics on the data path) -> a self-resetting ticket elects the last CTA, which folds all
 * slots into the pinned host-mapped gsb_kernel_out. Checksums are XOR / wrapping-add, so the
 * result does not depend on CTA scheduling order.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "gsb_internal.h"
#include "gsb_pattern.h"

#ifndef GSB_LAB
#define GSB_LAB 0
#endif

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxWarps = 16;
constexpr unsigned long long kNoBad = ~0ull;

struct Acc {
  uint32_t mm_words = 0, mm_bits = 0, cxor = 0, csum = 0, words = 0;
  unsigned long long first_bad = kNoBad;
};

// ---------------------------------------------------------------- memory-op helpers (PTX)

__device__ __forceinline__ uint4 ld_stream(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_stream(uint4 *p, const uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
// cache-operator flavours of the direct path (experiment knob GSB_DIRECT_FLAVOR):
// 0 = ld.cs/st.cs (streaming both ways), 1 = default ld/st, 2 = ld.L1::no_allocate + default st,
// 3 = ld.cg/st.cg (L2 only)
template <int F>
__device__ __forceinline__ uint4 ld_flavor(const uint4 *p) {
  uint4 v;
  if (F == 0) {
    return ld_stream(p);
  } else if (F == 1) {
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  } else if (F == 2) {
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  } else {
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  }
  return v;
}
// A refill stores to the address it has just loaded from, and the stored value (the next generation's pattern) does
// not depend on the loaded one. Left alone, the four stores of a thread issue right behind its four loads, while
// those loads are still in flight, and the memory system then orders every same-line load/store pair one at a time:
// 0.12-0.16 of the HBM copy peak in rounds 1 and 2 (profiles/sweep_r01_variants.json, sweep_r02.json "direct refill")
// against 0.93-0.97 for VERIFY or FILL alone. tools/lab/direct_anomaly.cu isolates it (profiles/direct_anomaly_r02.json):
// the same in-place walk runs at 6.1 TB/s when the stored value really depends on the loaded one or when the store
// trails its load by a tile, and at 1.8-3.0 TB/s when it does not. Round 1's remedy — naming the loaded register as
// an input of an EMPTY inline-asm statement — did nothing: a comment is not an instruction, so the dependency was gone
// before ptxas scheduled the store. A real one costs two integer ops per word: the loaded lanes are AND-ed with a
// kernel argument that is always 0 (the compiler cannot know) and XOR-ed into the value to store.
__device__ __forceinline__ uint4 after_load(uint4 nw, uint32_t loaded, uint32_t opaque_zero) {
  nw.x ^= loaded & opaque_zero;
  return nw;
}
template <int F>
__device__ __forceinline__ void st_flavor(uint4 *p, const uint4 v) {
  if (F == 0) {
    st_stream(p, v);
  } else if (F == 3) {
    asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  } else {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
  }
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
// TMA 1-D bulk copy shared -> global, tracked by the thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem),
               "r"(bytes)
               : "memory");
}
// same, carrying an L2 cache policy (createpolicy ... evict_first: streamed once, do not keep)
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar,
                                              uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          dst_smem),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void *dst, uint32_t src_smem, uint32_t bytes, uint64_t policy) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst),
               "r"(src_smem), "r"(bytes), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- pattern + compare

__device__ __forceinline__ uint4 pattern_from_key(uint32_t m, uint32_t seed_key) {
  uint4 p;
  p.x = (m * GSB_PAT_K0 + GSB_PAT_C0) ^ seed_key;
  p.y = (m * GSB_PAT_K1 + GSB_PAT_C1) ^ seed_key;
  p.z = (m * GSB_PAT_K2 + GSB_PAT_C2) ^ seed_key;
  p.w = (m * GSB_PAT_K3 + GSB_PAT_C3) ^ seed_key;
  return p;
}

__device__ __forceinline__ void fold_checksum(const uint4 v, Acc &acc) {
  acc.cxor ^= v.x ^ v.y ^ v.z ^ v.w;
  acc.csum += v.x + v.y + v.z + v.w;
}

__device__ __forceinline__ void compare_word(const uint4 v, const uint4 e, unsigned long long w, Acc &acc) {
  const uint32_t dx = v.x ^ e.x, dy = v.y ^ e.y, dz = v.z ^ e.z, dw = v.w ^ e.w;
  if ((dx | dy | dz | dw) != 0u) {  // cold path: a healthy device never takes it
    acc.mm_words += 1u;
    acc.mm_bits += __popc(dx) + __popc(dy) + __popc(dz) + __popc(dw);
    acc.first_bad = w < acc.first_bad ? w : acc.first_bad;
  }
}

// expected-seed key of the tile starting at absolute word w (tiles never straddle a granule when a
// table is in use: the shim enforces 64 KiB window alignment and granules are 64 MiB)
__device__ __forceinline__ uint32_t expect_key_of(const gsb_kernel_args &a, unsigned long long w) {
  const uint32_t s = a.seed_table ? __ldg(a.seed_table + (w >> a.granule_shift)) : a.seed_expect;
  return gsb_seed_key(s);
}

// one word: OP is a compile-time GSB_OP_*; `v` is the loaded value (ignored for FILL); returns the
// value to store (undefined for VERIFY)
template <int OP>
__device__ __forceinline__ uint4 process_word(const uint4 v, unsigned long long w, uint32_t key_expect,
                                              uint32_t key_write, Acc &acc) {
  const uint32_t m = gsb_word_key(w);
  uint4 nw = make_uint4(0, 0, 0, 0);
  if (OP != GSB_OP_FILL) {
    fold_checksum(v, acc);
    compare_word(v, pattern_from_key(m, key_expect), w, acc);
  }
  if (OP != GSB_OP_VERIFY) {
    nw = pattern_from_key(m, key_write);
    if (OP == GSB_OP_FILL) fold_checksum(nw, acc);
  }
  acc.words += 1u;
  return nw;
}

// ---------------------------------------------------------------- CTA epilogue

__device__ __forceinline__ void warp_fold(gsb_partial &p) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    p.mismatch_words += __shfl_xor_sync(0xffffffffu, p.mismatch_words, off);
    p.mismatch_bits += __shfl_xor_sync(0xffffffffu, p.mismatch_bits, off);
    p.words += __shfl_xor_sync(0xffffffffu, p.words, off);
    const unsigned long long o = __shfl_xor_sync(0xffffffffu, p.first_bad_word, off);
    p.first_bad_word = o < p.first_bad_word ? o : p.first_bad_word;
    p.checksum_xor ^= __shfl_xor_sync(0xffffffffu, p.checksum_xor, off);
    p.checksum_sum += __shfl_xor_sync(0xffffffffu, p.checksum_sum, off);
  }
}

__device__ __forceinline__ void merge(gsb_partial &d, const gsb_partial &s) {
  d.mismatch_words += s.mismatch_words;
  d.mismatch_bits += s.mismatch_bits;
  d.words += s.words;
  d.first_bad_word = s.first_bad_word < d.first_bad_word ? s.first_bad_word : d.first_bad_word;
  d.checksum_xor ^= s.checksum_xor;
  d.checksum_sum += s.checksum_sum;
}

__device__ __forceinline__ gsb_partial block_fold(gsb_partial p, gsb_partial *wslots) {
  warp_fold(p);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_warps = (int)(blockDim.x >> 5);
  if (lane == 0) wslots[warp] = p;
  __syncthreads();
  gsb_partial r = wslots[0];
  if (threadIdx.x == 0) {
    for (int i = 1; i < n_warps; i++) merge(r, wslots[i]);
  }
  return r;  // valid in thread 0
}

__device__ __forceinline__ gsb_partial ld_partial_cg(const gsb_partial *p) {
  gsb_partial r;
  r.mismatch_words = __ldcg(&p->mismatch_words);
  r.mismatch_bits = __ldcg(&p->mismatch_bits);
  r.first_bad_word = __ldcg(&p->first_bad_word);
  r.words = __ldcg(&p->words);
  r.checksum_xor = __ldcg(&p->checksum_xor);
  r.checksum_sum = __ldcg(&p->checksum_sum);
  return r;
}

__device__ void finish(const gsb_kernel_args &a, const Acc &acc) {
  __shared__ gsb_partial wslots[kMaxWarps];
  __shared__ int is_last;
  gsb_partial p;
  p.mismatch_words = acc.mm_words;
  p.mismatch_bits = acc.mm_bits;
  p.first_bad_word = acc.first_bad;
  p.words = acc.words;
  p.checksum_xor = acc.cxor;
  p.checksum_sum = acc.csum;
  gsb_partial cta = block_fold(p, wslots);
  if (threadIdx.x == 0) {
    a.partials[blockIdx.x] = cta;
    __threadfence();
    const unsigned prev = atomicAdd(a.ticket, 1u);
    is_last = (prev == gridDim.x - 1u);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  gsb_partial t;
  t.mismatch_words = 0;
  t.mismatch_bits = 0;
  t.first_bad_word = kNoBad;
  t.words = 0;
  t.checksum_xor = 0;
  t.checksum_sum = 0;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += blockDim.x) merge(t, ld_partial_cg(a.partials + i));
  __syncthreads();  // wslots reuse
  gsb_partial all = block_fold(t, wslots);
  if (a.table_update) {
    // every other CTA has finished: record the generation now held by each granule this window
    // covered completely (plus the arena's ragged tail granule when the window reaches the end)
    const unsigned long long gw = 1ull << a.granule_shift;
    const unsigned long long end = a.first_word + a.n_words;
    const unsigned long long g0 = (a.first_word + gw - 1) >> a.granule_shift;
    unsigned long long g1 = end >> a.granule_shift;
    if (end == a.arena_words && (a.arena_words & (gw - 1))) g1 += 1;
    for (unsigned long long g = g0 + threadIdx.x; g < g1; g += blockDim.x) a.table_update[g] = a.seed_write;
  }
  if (threadIdx.x == 0) {
    *a.ticket = 0u;  // self-reset: the next launch on this stream starts from 0
    if (a.tile_counter) *a.tile_counter = 0ull;
    gsb_kernel_out *o = a.out;
    o->mismatch_words = all.mismatch_words;
    o->mismatch_bits = all.mismatch_bits;
    o->first_bad_word = all.first_bad_word;
    o->checksum_xor = all.checksum_xor;
    o->checksum_sum = all.checksum_sum;
    o->words_done = all.words;
    __threadfence_system();
    o->done_flag = a.launch_seq;
  }
}

// ---------------------------------------------------------------- DIRECT

template <int OP, int U, int F>
__global__ void __launch_bounds__(kThreads) probe_direct(const gsb_kernel_args a) {
  constexpr unsigned long long TILE = (unsigned long long)kThreads * U;
  const unsigned long long n_tiles = (a.n_words + TILE - 1) / TILE;
  const uint32_t key_write = gsb_seed_key(a.seed_write);
  uint4 *__restrict__ win = a.base + a.first_word;
  Acc acc;
  for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    const unsigned long long l0 = t * TILE + threadIdx.x;
    const uint32_t key_expect = (OP != GSB_OP_FILL) ? expect_key_of(a, a.first_word + t * TILE) : 0u;
    uint4 v[U];
    if ((t + 1) * TILE <= a.n_words) {
      if (OP != GSB_OP_FILL) {
#pragma unroll
        for (int u = 0; u < U; u++) v[u] = ld_flavor<F>(win + l0 + (unsigned long long)u * kThreads);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const unsigned long long l = l0 + (unsigned long long)u * kThreads;
        uint4 nw = process_word<OP>(v[u], a.first_word + l, key_expect, key_write, acc);
        if (OP == GSB_OP_VERIFY_REFILL) nw = after_load(nw, v[u].x ^ v[u].y ^ v[u].z ^ v[u].w, a.opaque_zero);
        if (OP != GSB_OP_VERIFY) st_flavor<F>(win + l, nw);
      }
    } else {  // ragged last tile
#pragma unroll
      for (int u = 0; u < U; u++) {
        const unsigned long long l = l0 + (unsigned long long)u * kThreads;
        if (l < a.n_words) {
          if (OP != GSB_OP_FILL) v[u] = ld_flavor<F>(win + l);
          uint4 nw = process_word<OP>(v[u], a.first_word + l, key_expect, key_write, acc);
          if (OP == GSB_OP_VERIFY_REFILL) nw = after_load(nw, v[u].x ^ v[u].y ^ v[u].z ^ v[u].w, a.opaque_zero);
          if (OP != GSB_OP_VERIFY) st_flavor<F>(win + l, nw);
        }
      }
    }
  }
  finish(a, acc);
}

// ---------------------------------------------------------------- CPASYNC (LDGSTS ring)

template <int OP, int U, int S>
__global__ void __launch_bounds__(kThreads) probe_cpasync(const gsb_kernel_args a) {
  static_assert(OP != GSB_OP_FILL, "FILL has no loads to stage");
  constexpr unsigned long long TILE = (unsigned long long)kThreads * U;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint4 *ring = reinterpret_cast<uint4 *>(smem_raw);  // [S][TILE]
  const unsigned long long n_tiles = (a.n_words + TILE - 1) / TILE;
  const unsigned long long my_n =
      n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0ull;
  const uint32_t key_write = gsb_seed_key(a.seed_write);
  uint4 *__restrict__ win = a.base + a.first_word;
  Acc acc;

  auto issue = [&](unsigned long long k) {
    if (k < my_n) {
      const unsigned long long t = blockIdx.x + k * gridDim.x;
      const int s = (int)(k % S);
#pragma unroll
      for (int u = 0; u < U; u++) {
        const unsigned long long l = t * TILE + (unsigned long long)u * kThreads + threadIdx.x;
        if (l < a.n_words) cp_async16(smem_u32(ring + s * TILE + u * kThreads + threadIdx.x), win + l);
      }
    }
    cp_async_commit();  // always commit: keeps the group count uniform
  };

#pragma unroll
  for (int k = 0; k < S - 1; k++) issue(k);
  for (unsigned long long k = 0; k < my_n; k++) {
    issue(k + S - 1);
    cp_async_wait<S - 1>();  // tile k's copies (this thread's own 16 B slots) have landed
    const unsigned long long t = blockIdx.x + k * gridDim.x;
    const int s = (int)(k % S);
    const uint32_t key_expect = expect_key_of(a, a.first_word + t * TILE);
#pragma unroll
    for (int u = 0; u < U; u++) {
      const unsigned long long l = t * TILE + (unsigned long long)u * kThreads + threadIdx.x;
      if (l < a.n_words) {
        // each thread reads back only the slots it copied itself: no CTA barrier needed
        const uint4 v = ring[s * TILE + u * kThreads + threadIdx.x];
        const uint4 nw = process_word<OP>(v, a.first_word + l, key_expect, key_write, acc);
        if (OP != GSB_OP_VERIFY) st_stream(win + l, nw);
      }
    }
  }
  cp_async_wait<0>();
  finish(a, acc);
}

// ---------------------------------------------------------------- BULK (TMA 1-D ring)

template <int OP, int U, int S>
__global__ void __launch_bounds__(kThreads) probe_bulk(const gsb_kernel_args a) {
  constexpr unsigned long long TILE = (unsigned long long)kThreads * U;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint4 *ring = reinterpret_cast<uint4 *>(smem_raw);  // [S][TILE]
  __shared__ __align__(8) unsigned long long full_bar[S];
  const unsigned long long n_tiles = (a.n_words + TILE - 1) / TILE;
  const unsigned long long my_n =
      n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0ull;
  const uint32_t key_write = gsb_seed_key(a.seed_write);
  uint4 *__restrict__ win = a.base + a.first_word;
  const bool leader = threadIdx.x == 0;
  const uint64_t policy = a.l2_hint ? l2_evict_first_policy() : 0ull;
  Acc acc;

  auto tile_words = [&](unsigned long long k) -> uint32_t {
    const unsigned long long t = blockIdx.x + k * gridDim.x;
    const unsigned long long left = a.n_words - t * TILE;
    return (uint32_t)(left < TILE ? left : TILE);
  };
  auto load = [&](unsigned long long k) {  // leader only
    const unsigned long long t = blockIdx.x + k * gridDim.x;
    const int s = (int)(k % S);
    const uint32_t bytes = tile_words(k) * 16u;
    mbar_expect_tx(smem_u32(&full_bar[s]), bytes);
    if (a.l2_hint & 1u)
      bulk_g2s_hint(smem_u32(ring + s * TILE), win + t * TILE, bytes, smem_u32(&full_bar[s]), policy);
    else
      bulk_g2s(smem_u32(ring + s * TILE), win + t * TILE, bytes, smem_u32(&full_bar[s]));
  };

  if (leader) {
#pragma unroll
    for (int s = 0; s < S; s++) mbar_init(smem_u32(&full_bar[s]), 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // prologue: VERIFY can keep all S stages in flight (a consumed stage is refilled at once);
  // VERIFY_REFILL keeps S-1 (a stage is refilled one iteration after its store was issued)
  constexpr int kPrologue = (OP == GSB_OP_VERIFY) ? S : S - 1;
  if (OP != GSB_OP_FILL && leader) {
    for (int k = 0; k < kPrologue; k++)
      if ((unsigned long long)k < my_n) load(k);
  }

  for (unsigned long long k = 0; k < my_n; k++) {
    const unsigned long long t = blockIdx.x + k * gridDim.x;
    const int s = (int)(k % S);
    const uint32_t nw_tile = tile_words(k);
    uint4 *stage = ring + s * TILE;
    if (OP == GSB_OP_FILL) {
      // the bulk store that last read this stage (tile k-S) must have drained it
      if (leader) bulk_wait_read<S - 1>();
      __syncthreads();
    } else {
      mbar_wait(smem_u32(&full_bar[s]), (uint32_t)((k / S) & 1ull));
    }
    const uint32_t key_expect = (OP != GSB_OP_FILL) ? expect_key_of(a, a.first_word + t * TILE) : 0u;
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t i = u * kThreads + threadIdx.x;
      if (i < nw_tile) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (OP != GSB_OP_FILL) v = stage[i];
        const uint4 nw = process_word<OP>(v, a.first_word + t * TILE + i, key_expect, key_write, acc);
        if (OP != GSB_OP_VERIFY) stage[i] = nw;
      }
    }
    if (OP != GSB_OP_VERIFY) fence_proxy_async_smem();  // generic-proxy writes -> visible to the bulk engine
    __syncthreads();  // every thread is done with this stage (reads, and writes + proxy fence)
    if (leader) {
      if (OP == GSB_OP_VERIFY) {
        if (k + S < my_n) load(k + S);  // same stage, just released by the barrier
      } else {
        if (a.l2_hint & 2u)
          bulk_s2g_hint(win + t * TILE, smem_u32(stage), nw_tile * 16u, policy);
        else
          bulk_s2g(win + t * TILE, smem_u32(stage), nw_tile * 16u);
        bulk_commit();
        if (OP == GSB_OP_VERIFY_REFILL && k + S - 1 < my_n) {
          // refill stage (k-1)%S: its store (tile k-1) must have finished READING shared memory;
          // one group (tile k's store, just committed) may stay in flight
          if (k >= 1) bulk_wait_read<1>();
          load(k + S - 1);
        }
      }
    }
  }
  if (leader && OP != GSB_OP_VERIFY) bulk_wait_all<0>();  // shared memory must outlive the stores
  finish(a, acc);
}

// ---------------------------------------------------------------- BULKD (TMA ring, dynamic tile scheduler)
//
// BULK with the static stride (tile = blockIdx + k*grid) replaced by an atomic tile counter: the leader
// claims the next tile when it issues that tile's load, so CTAs that run ahead simply take more tiles and
// the kernel ends with every CTA busy (no straggler tail). Claims are made in time order, so concurrently
// processed tiles stay neighbours in memory. The claimed index travels to the consumers through shared
// memory under the stage's mbarrier (arrive = release, wait = acquire).
template <int OP, int U, int S>
__global__ void __launch_bounds__(kThreads) probe_bulk_dyn(const gsb_kernel_args a) {
  constexpr unsigned long long TILE = (unsigned long long)kThreads * U;
  constexpr unsigned long long kDone = ~0ull;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint4 *ring = reinterpret_cast<uint4 *>(smem_raw);  // [S][TILE]
  __shared__ __align__(8) unsigned long long full_bar[S];
  __shared__ unsigned long long tile_of[S];
  const unsigned long long n_tiles = (a.n_words + TILE - 1) / TILE;
  const uint32_t key_write = gsb_seed_key(a.seed_write);
  uint4 *__restrict__ win = a.base + a.first_word;
  const bool leader = threadIdx.x == 0;
  Acc acc;
  // the leader always holds one claim in hand: the atomic for the NEXT tile is issued when the current one
  // is consumed, so its ~L2 round trip overlaps the tile being processed instead of delaying the next load
  unsigned long long in_hand = 0;
  if (leader) in_hand = atomicAdd(a.tile_counter, 1ull);
  auto claim = [&]() -> unsigned long long {
    const unsigned long long t = in_hand;
    in_hand = atomicAdd(a.tile_counter, 1ull);
    return t < n_tiles ? t : kDone;
  };
  auto words_of = [&](unsigned long long t) -> uint32_t {
    const unsigned long long left = a.n_words - t * TILE;
    return (uint32_t)(left < TILE ? left : TILE);
  };
  auto load_next = [&](unsigned long long k) {  // leader only: claim a tile for slot k and start its load
    const int s = (int)(k % S);
    const unsigned long long t = claim();
    tile_of[s] = t;
    if (t != kDone) {
      const uint32_t bytes = words_of(t) * 16u;
      mbar_expect_tx(smem_u32(&full_bar[s]), bytes);
      bulk_g2s(smem_u32(ring + s * TILE), win + t * TILE, bytes, smem_u32(&full_bar[s]));
    } else {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full_bar[s])) : "memory");
    }
  };

  if (leader) {
#pragma unroll
    for (int s = 0; s < S; s++) mbar_init(smem_u32(&full_bar[s]), 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  constexpr int kPrologue = (OP == GSB_OP_VERIFY) ? S : S - 1;
  if (OP != GSB_OP_FILL && leader)
    for (int k = 0; k < kPrologue; k++) load_next(k);

  for (unsigned long long k = 0;; k++) {
    const int s = (int)(k % S);
    if (OP == GSB_OP_FILL) {
      if (leader) {
        bulk_wait_read<S - 1>();  // the store that last read this stage (slot k-S) has drained it
        tile_of[s] = claim();
      }
      __syncthreads();
    } else {
      mbar_wait(smem_u32(&full_bar[s]), (uint32_t)((k / S) & 1ull));
    }
    const unsigned long long t = tile_of[s];
    if (t == kDone) break;  // uniform: every thread reads the same slot after the same barrier
    const uint32_t nw_tile = words_of(t);
    uint4 *stage = ring + s * TILE;
    const uint32_t key_expect = (OP != GSB_OP_FILL) ? expect_key_of(a, a.first_word + t * TILE) : 0u;
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t i = u * kThreads + threadIdx.x;
      if (i < nw_tile) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (OP != GSB_OP_FILL) v = stage[i];
        const uint4 nw = process_word<OP>(v, a.first_word + t * TILE + i, key_expect, key_write, acc);
        if (OP != GSB_OP_VERIFY) stage[i] = nw;
      }
    }
    if (OP != GSB_OP_VERIFY) fence_proxy_async_smem();
    __syncthreads();
    if (leader) {
      if (OP == GSB_OP_VERIFY) {
        load_next(k + S);
      } else {
        bulk_s2g(win + t * TILE, smem_u32(stage), nw_tile * 16u);
        bulk_commit();
        if (OP == GSB_OP_VERIFY_REFILL) {
          if (k >= 1) bulk_wait_read<1>();  // stage (k-1)%S: its store has finished reading shared memory
          load_next(k + S - 1);
        }
      }
    }
  }
  if (leader && OP != GSB_OP_VERIFY) bulk_wait_all<0>();
  finish(a, acc);
}

// ---------------------------------------------------------------- BULKW (per-warp TMA rings)
//
// Same data path as BULK, but every warp owns a private ring of S stages of WPL*512 bytes and its own
// mbarriers; lane 0 issues that warp's bulk loads/stores. Nothing on the data path synchronises the
// CTA (only __syncwarp), so a warp waiting on HBM never holds the other warps at a barrier. The warps
// of a CTA take adjacent chunks of one contiguous super-tile, which keeps DRAM pages hot.
template <int OP, int WPL, int S, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) probe_bulk_warp(const gsb_kernel_args a) {
  constexpr unsigned long long TILE = (unsigned long long)WPL * 32;  // words per warp-tile
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint4 *ring = reinterpret_cast<uint4 *>(smem_raw) + (size_t)warp * S * TILE;  // [S][TILE]
  __shared__ __align__(8) unsigned long long full_bar[WARPS][S];
  const unsigned long long n_tiles = (a.n_words + TILE - 1) / TILE;
  // k-th tile of this warp: warps of one CTA sit side by side inside a WARPS*TILE super-tile
  auto tile_of = [&](unsigned long long k) -> unsigned long long {
    return (blockIdx.x + k * gridDim.x) * WARPS + warp;
  };
  const unsigned long long first = tile_of(0);
  const unsigned long long stride = (unsigned long long)gridDim.x * WARPS;
  const unsigned long long my_n = n_tiles > first ? (n_tiles - first + stride - 1) / stride : 0ull;
  const uint32_t key_write = gsb_seed_key(a.seed_write);
  uint4 *__restrict__ win = a.base + a.first_word;
  const bool leader = lane == 0;
  const uint64_t policy = a.l2_hint ? l2_evict_first_policy() : 0ull;
  Acc acc;

  auto tile_words = [&](unsigned long long k) -> uint32_t {
    const unsigned long long left = a.n_words - tile_of(k) * TILE;
    return (uint32_t)(left < TILE ? left : TILE);
  };
  auto load = [&](unsigned long long k) {  // leader only
    const int s = (int)(k % S);
    const uint32_t bytes = tile_words(k) * 16u;
    mbar_expect_tx(smem_u32(&full_bar[warp][s]), bytes);
    if (a.l2_hint & 1u)
      bulk_g2s_hint(smem_u32(ring + s * TILE), win + tile_of(k) * TILE, bytes, smem_u32(&full_bar[warp][s]), policy);
    else
      bulk_g2s(smem_u32(ring + s * TILE), win + tile_of(k) * TILE, bytes, smem_u32(&full_bar[warp][s]));
  };

  if (leader) {
#pragma unroll
    for (int s = 0; s < S; s++) mbar_init(smem_u32(&full_bar[warp][s]), 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  constexpr int kPrologue = (OP == GSB_OP_VERIFY) ? S : S - 1;
  if (OP != GSB_OP_FILL && leader) {
    for (int k = 0; k < kPrologue; k++)
      if ((unsigned long long)k < my_n) load(k);
  }

  for (unsigned long long k = 0; k < my_n; k++) {
    const unsigned long long t = tile_of(k);
    const int s = (int)(k % S);
    const uint32_t nw_tile = tile_words(k);
    uint4 *stage = ring + s * TILE;
    if (OP == GSB_OP_FILL) {
      if (leader) bulk_wait_read<S - 1>();  // the store that last read this stage has drained it
      __syncwarp();
    } else {
      mbar_wait(smem_u32(&full_bar[warp][s]), (uint32_t)((k / S) & 1ull));
    }
    const uint32_t key_expect = (OP != GSB_OP_FILL) ? expect_key_of(a, a.first_word + t * TILE) : 0u;
    if (nw_tile == TILE) {
      uint4 v[WPL];
      if (OP != GSB_OP_FILL) {
#pragma unroll
        for (int u = 0; u < WPL; u++) v[u] = stage[u * 32 + lane];
      }
#pragma unroll
      for (int u = 0; u < WPL; u++) {
        const uint4 nw = process_word<OP>(v[u], a.first_word + t * TILE + u * 32 + lane, key_expect, key_write, acc);
        if (OP != GSB_OP_VERIFY) stage[u * 32 + lane] = nw;
      }
    } else {
#pragma unroll
      for (int u = 0; u < WPL; u++) {
        const uint32_t i = u * 32 + lane;
        if (i < nw_tile) {
          uint4 v = make_uint4(0, 0, 0, 0);
          if (OP != GSB_OP_FILL) v = stage[i];
          const uint4 nw = process_word<OP>(v, a.first_word + t * TILE + i, key_expect, key_write, acc);
          if (OP != GSB_OP_VERIFY) stage[i] = nw;
        }
      }
    }
    if (OP != GSB_OP_VERIFY) fence_proxy_async_smem();
    __syncwarp();  // the whole warp is done with this stage
    if (leader) {
      if (OP == GSB_OP_VERIFY) {
        if (k + S < my_n) load(k + S);
      } else {
        if (a.l2_hint & 2u)
          bulk_s2g_hint(win + t * TILE, smem_u32(stage), nw_tile * 16u, policy);
        else
          bulk_s2g(win + t * TILE, smem_u32(stage), nw_tile * 16u);
        bulk_commit();
        if (OP == GSB_OP_VERIFY_REFILL && k + S - 1 < my_n) {
          if (k >= 1) bulk_wait_read<1>();
          load(k + S - 1);
        }
      }
    }
  }
  if (leader && OP != GSB_OP_VERIFY) bulk_wait_all<0>();
  finish(a, acc);
}

// ---------------------------------------------------------------- geometry + dispatch

constexpr int kDirectU = 4;
constexpr int kCpU = 4, kCpS = 4;      // 16 KiB tiles x 4 stages = 64 KiB / CTA
constexpr int kBulkU = 4, kBulkS = 4;  // 16 KiB tiles x 4 stages = 64 KiB / CTA
constexpr uint32_t kCpSmem = kCpU * kThreads * 16 * kCpS;
constexpr int kMaxCtasPerSm = 8;

template <typename K>
int resident_ctas(K kernel, uint32_t smem, int threads) {
  int n = 0;
  if (smem > 32 * 1024) {  // dynamic + the kernel's static shared memory may cross the 48 KiB default
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
      return -1;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess) return -1;
  return n;
}

typedef void (*probe_fn)(const gsb_kernel_args);

// knob GSB_DYN_FILL=0: FILL under GSB_VARIANT_BULKD falls back to the static kernel (32 KiB x 6)
bool dyn_fill() {
#if GSB_LAB
  static const bool f = [] {
    const char *e = getenv("GSB_DYN_FILL");
    return !e || atoi(e) != 0;
  }();
  return f;
#else
  return true;
#endif
}

// experiment knob GSB_BULKW_CFG: warp-tile size x ring depth x warps per CTA of the per-warp TMA path
int bulkw_cfg() {
#if !GSB_LAB
  return 0;
#endif
  static const int f = [] {
    const char *e = getenv("GSB_BULKW_CFG");
    const int v = e ? atoi(e) : 0;
    return v < 0 || v > 7 ? 0 : v;
  }();
  return f;
}

// experiment knob GSB_BULK_CFG: tile size x ring depth of the TMA path (unset = per-op default)
int bulk_cfg() {
#if !GSB_LAB
  return -1;
#endif
  static const int f = [] {
    const char *e = getenv("GSB_BULK_CFG");
    const int v = e ? atoi(e) : -1;
    return v < 0 || v > 8 ? -1 : v;
  }();
  return f;
}

// experiment knob GSB_L2_HINT: bit0 loads / bit1 stores of the TMA paths carry L2::evict_first
uint32_t l2_hint_knob() {
#if !GSB_LAB
  return 0u;
#endif
  static const uint32_t f = [] {
    const char *e = getenv("GSB_L2_HINT");
    return e ? (uint32_t)atoi(e) & 3u : 0u;
  }();
  return f;
}

int direct_flavor() {
#if !GSB_LAB
  return 0;
#endif
  static const int f = [] {
    const char *e = getenv("GSB_DIRECT_FLAVOR");
    const int v = e ? atoi(e) : 0;
    return v < 0 || v > 3 ? 0 : v;
  }();
  return f;
}

probe_fn pick(uint32_t op, uint32_t variant, uint32_t *smem, uint32_t *threads) {
  *smem = 0;
  *threads = kThreads;
  switch (variant) {
    case GSB_VARIANT_DIRECT:
      switch (direct_flavor()) {
#define GSB_DIRECT_CASE(F)                                                                   \
  case F:                                                                                    \
    if (op == GSB_OP_FILL) return probe_direct<GSB_OP_FILL, kDirectU, F>;                    \
    if (op == GSB_OP_VERIFY) return probe_direct<GSB_OP_VERIFY, kDirectU, F>;                \
    if (op == GSB_OP_VERIFY_REFILL) return probe_direct<GSB_OP_VERIFY_REFILL, kDirectU, F>;  \
    return nullptr;
        GSB_DIRECT_CASE(0)
#if GSB_LAB
        GSB_DIRECT_CASE(1)
        GSB_DIRECT_CASE(2)
        GSB_DIRECT_CASE(3)
#endif
#undef GSB_DIRECT_CASE
      }
      return nullptr;
    case GSB_VARIANT_CPASYNC:
      // FILL issues no loads, so there is nothing to stage: it takes the direct store path
      if (op == GSB_OP_FILL) return probe_direct<GSB_OP_FILL, kDirectU, 0>;
      *smem = kCpSmem;
      if (op == GSB_OP_VERIFY) return probe_cpasync<GSB_OP_VERIFY, kCpU, kCpS>;
      if (op == GSB_OP_VERIFY_REFILL) return probe_cpasync<GSB_OP_VERIFY_REFILL, kCpU, kCpS>;
      return nullptr;
    case GSB_VARIANT_BULK:
      // shipped defaults (profiles/sweep_r01_knobs2.json): 32 KiB tiles; a pure-store FILL wants one fat
      // CTA per SM with a deep ring, the loading ops want 2 CTAs/SM x 3 stages
      switch (bulk_cfg() >= 0 ? bulk_cfg() : ((GSB_LAB && op == GSB_OP_FILL) ? 5 : 1)) {
#define GSB_BULK_CASE(ID, U, S)                                                              \
  case ID:                                                                                   \
    *smem = U * kThreads * 16 * S;                                                           \
    if (op == GSB_OP_FILL) return probe_bulk<GSB_OP_FILL, U, S>;                             \
    if (op == GSB_OP_VERIFY) return probe_bulk<GSB_OP_VERIFY, U, S>;                         \
    if (op == GSB_OP_VERIFY_REFILL) return probe_bulk<GSB_OP_VERIFY_REFILL, U, S>;           \
    return nullptr;
        GSB_BULK_CASE(1, 8, 3)            // 32 KiB x 3 = 96 KiB (2 CTAs/SM)   <- shipped
#if GSB_LAB
        GSB_BULK_CASE(0, kBulkU, kBulkS)  // 16 KiB x 4 stages = 64 KiB/CTA (3 CTAs/SM)
        GSB_BULK_CASE(2, 4, 6)            // 16 KiB x 6 = 96 KiB (2 CTAs/SM)
        GSB_BULK_CASE(3, 2, 8)            //  8 KiB x 8 = 64 KiB (3 CTAs/SM)
        GSB_BULK_CASE(4, 4, 3)            // 16 KiB x 3 = 48 KiB (4 CTAs/SM)
        GSB_BULK_CASE(5, 8, 6)            // 32 KiB x 6 = 192 KiB (1 CTA/SM)
        GSB_BULK_CASE(6, 16, 3)           // 64 KiB x 3 = 192 KiB (1 CTA/SM)
        GSB_BULK_CASE(7, 16, 2)           // 64 KiB x 2 = 128 KiB (1 CTA/SM)
        GSB_BULK_CASE(8, 8, 4)            // 32 KiB x 4 = 128 KiB (1 CTA/SM)
#endif
#undef GSB_BULK_CASE
      }
      return nullptr;
    case GSB_VARIANT_BULKD:
      // dynamic schedule for the loading ops; FILL keeps the static BULK kernel (32 KiB x 6, 1 CTA/SM)
#if GSB_LAB
      if (op == GSB_OP_FILL && !dyn_fill()) {
        *smem = 8 * kThreads * 16 * 6;
        return probe_bulk<GSB_OP_FILL, 8, 6>;
      }
#endif
      switch (bulk_cfg() >= 0 ? bulk_cfg() : 1) {  // 32 KiB x 3, 2 CTAs/SM for all three ops
#define GSB_BULKD_CASE(ID, U, S)                                                             \
  case ID:                                                                                   \
    *smem = U * kThreads * 16 * S;                                                           \
    if (op == GSB_OP_FILL) return probe_bulk_dyn<GSB_OP_FILL, U, S>;                         \
    if (op == GSB_OP_VERIFY) return probe_bulk_dyn<GSB_OP_VERIFY, U, S>;                     \
    return probe_bulk_dyn<GSB_OP_VERIFY_REFILL, U, S>;
        GSB_BULKD_CASE(1, 8, 3)  // <- shipped
#if GSB_LAB
        GSB_BULKD_CASE(0, 4, 4)
        GSB_BULKD_CASE(2, 4, 6)
        GSB_BULKD_CASE(3, 2, 8)
        GSB_BULKD_CASE(4, 4, 3)
        GSB_BULKD_CASE(5, 8, 6)
        GSB_BULKD_CASE(6, 16, 3)  // 64 KiB x 3 = 192 KiB (1 CTA/SM)
        GSB_BULKD_CASE(7, 16, 2)  // 64 KiB x 2 = 128 KiB (1 CTA/SM)
        GSB_BULKD_CASE(8, 8, 4)   // 32 KiB x 4 = 128 KiB (1 CTA/SM)
#endif
#undef GSB_BULKD_CASE
      }
      return nullptr;
    case GSB_VARIANT_BULKW:
      switch (bulkw_cfg()) {
#define GSB_BULKW_CASE(ID, WPL, S, WARPS)                                                    \
  case ID:                                                                                   \
    *smem = WPL * 512 * S * WARPS;                                                           \
    *threads = WARPS * 32;                                                                   \
    if (op == GSB_OP_FILL) return probe_bulk_warp<GSB_OP_FILL, WPL, S, WARPS>;               \
    if (op == GSB_OP_VERIFY) return probe_bulk_warp<GSB_OP_VERIFY, WPL, S, WARPS>;           \
    if (op == GSB_OP_VERIFY_REFILL) return probe_bulk_warp<GSB_OP_VERIFY_REFILL, WPL, S, WARPS>; \
    return nullptr;
        GSB_BULKW_CASE(0, 8, 3, 8)    //  4 KiB x 3 x 8 warps =  96 KiB/CTA (2 CTAs/SM, 16 warps/SM)
#if GSB_LAB
        GSB_BULKW_CASE(1, 8, 6, 4)    //  4 KiB x 6 x 4 warps =  96 KiB     (2 CTAs/SM,  8 warps/SM)
        GSB_BULKW_CASE(2, 16, 3, 4)   //  8 KiB x 3 x 4 warps =  96 KiB     (2 CTAs/SM,  8 warps/SM)
        GSB_BULKW_CASE(3, 4, 6, 8)    //  2 KiB x 6 x 8 warps =  96 KiB     (2 CTAs/SM, 16 warps/SM)
        GSB_BULKW_CASE(4, 8, 4, 4)    //  4 KiB x 4 x 4 warps =  64 KiB     (3 CTAs/SM, 12 warps/SM)
        GSB_BULKW_CASE(5, 16, 3, 8)   //  8 KiB x 3 x 8 warps = 192 KiB     (1 CTA/SM,   8 warps/SM)
        GSB_BULKW_CASE(6, 8, 3, 16)   //  4 KiB x 3 x 16 warps = 192 KiB    (1 CTA/SM,  16 warps/SM)
        GSB_BULKW_CASE(7, 8, 2, 8)    //  4 KiB x 2 x 8 warps =  64 KiB     (3 CTAs/SM, 24 warps/SM)
#endif
#undef GSB_BULKW_CASE
      }
      return nullptr;
    default:
      return nullptr;
  }
}

// test hook (gsb_test_stall): hold the stream for `ns` nanoseconds
__global__ void stall_kernel(unsigned long long ns) {
  unsigned long long t0, t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  do {
    __nanosleep(1000);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  } while (t - t0 < ns);
}

}  // namespace

int gsb_kernel_stall(unsigned long long ns, cudaStream_t stream) {
  stall_kernel<<<1, 1, 0, stream>>>(ns);
  return (int)cudaGetLastError();
}

uint32_t gsb_kernel_max_grid(int sm_count) { return (uint32_t)(sm_count * kMaxCtasPerSm); }

int gsb_kernel_geometry(uint32_t op, uint32_t variant, uint32_t grid_request, int sm_count,
                        gsb_launch_geom *geom) {
  // shipped choice per op (profiles/sweep_r01_dynamic_scheduler{,2}.json): the writing ops want the dynamic
  // tile scheduler (refill 1.07 vs 0.95, fill 1.07 vs 1.02 of the copy peak on the full arena); VERIFY is as
  // fast or faster on the static one
  if (variant == GSB_VARIANT_AUTO) variant = op == GSB_OP_VERIFY ? GSB_VARIANT_BULK : GSB_VARIANT_BULKD;
  uint32_t smem = 0, threads = 0;
  probe_fn fn = pick(op, variant, &smem, &threads);
  if (!fn) return (int)cudaErrorInvalidValue;
  int per_sm = resident_ctas(fn, smem, (int)threads);
  if (per_sm <= 0) return (int)cudaErrorInvalidDeviceFunction;
  if (per_sm > kMaxCtasPerSm) per_sm = kMaxCtasPerSm;
  uint32_t grid = (uint32_t)(per_sm * sm_count);  // one full wave of resident CTAs: persistent
  if (grid_request) grid = grid_request < gsb_kernel_max_grid(sm_count) ? grid_request : gsb_kernel_max_grid(sm_count);
  geom->variant = variant;
  geom->grid = grid;
  geom->block = threads;
  geom->smem_bytes = smem;
  return 0;
}

int gsb_kernel_launch(uint32_t op, const gsb_launch_geom *geom, const gsb_kernel_args *args,
                      cudaStream_t stream) {
  uint32_t smem = 0, threads = 0;
  probe_fn fn = pick(op, geom->variant, &smem, &threads);
  if (!fn) return (int)cudaErrorInvalidValue;
  gsb_kernel_args a = *args;
  a.l2_hint = l2_hint_knob();
  a.opaque_zero = 0u;
  fn<<<geom->grid, threads, smem, stream>>>(a);
  return (int)cudaGetLastError();
}
