// direct_anomaly.cu — isolates why the DIRECT control variant (ld.global.v4 -> registers -> st.global.v4 to the SAME
// address) ran VERIFY_REFILL at 0.12-0.15 of the HBM copy peak in round 1 while VERIFY alone ran at 1.03 and FILL at
// 0.91 (DESIGN.md §5 "open anomaly"). Standalone: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/direct_anomaly
// tools/lab/direct_anomaly.cu ; prints one JSON line. Not part of the product.
//
//   A inplace          load tile k, store tile k (what DIRECT does)
//   B copy             load from X, store the same value pattern to Y = X + W (different lines): a plain copy
//   C inplace_lag1     software-pipelined in registers: load tile k+1 BEFORE storing tile k (store trails its load by a tile)
//   D inplace_far      each CTA stores tile k of a DIFFERENT region than it loads (loads region 0, stores region 1 of the
//                      same buffer, both walked once): in-place traffic mix, no same-line read->write from one SM
//   E inplace_nofeed   like A but the stored value does not depend on the load and no ordering is forced (the r01 start)
//   F inplace_u1/u8    A with 1 / 8 independent 16 B loads in flight per thread
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\":\"%s at %s:%d\"}\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint4 ldg(const uint4 *p) {
  uint4 v;
  asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stg(uint4 *p, uint4 v) {
  asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 mix(uint4 v, uint32_t k) {
  v.x = v.x * 0x9E3779B1u + k; v.y ^= v.x >> 7; v.z += v.y * 0x85EBCA6Bu; v.w ^= v.z + k;
  return v;
}

template <int U, int MODE>
__global__ void __launch_bounds__(256) k(uint4 *base, unsigned long long n_words, unsigned long long store_off, uint32_t seed, uint32_t *sink) {
  const unsigned long long TILE = 256ull * U;
  const unsigned long long n_tiles = n_words / TILE;
  uint32_t acc = 0;
  if (MODE == 2) {  // lag-1 software pipeline
    uint4 cur[U], nxt[U];
    unsigned long long t = blockIdx.x;
    if (t < n_tiles) {
#pragma unroll
      for (int u = 0; u < U; u++) cur[u] = ldg(base + t * TILE + u * 256 + threadIdx.x);
    }
    for (; t < n_tiles; t += gridDim.x) {
      const unsigned long long tn = t + gridDim.x;
      if (tn < n_tiles) {
#pragma unroll
        for (int u = 0; u < U; u++) nxt[u] = ldg(base + tn * TILE + u * 256 + threadIdx.x);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        acc ^= cur[u].x ^ cur[u].w;
        stg(base + t * TILE + u * 256 + threadIdx.x, mix(cur[u], seed));
      }
#pragma unroll
      for (int u = 0; u < U; u++) cur[u] = nxt[u];
    }
  } else {
    for (unsigned long long t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; u++) v[u] = ldg(base + t * TILE + u * 256 + threadIdx.x);
#pragma unroll
      for (int u = 0; u < U; u++) {
        acc ^= v[u].x ^ v[u].w;
        uint4 w = MODE == 4 ? make_uint4(seed + u, seed ^ (uint32_t)t, threadIdx.x, 7u) : mix(v[u], seed);
        stg(base + store_off + t * TILE + u * 256 + threadIdx.x, w);
      }
    }
  }
  if (acc == 0x12345u) *sink = acc;
}

template <int U, int MODE>
float run(uint4 *buf, unsigned long long n_words, unsigned long long store_off, int grid, uint32_t *sink, int reps) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<U, MODE><<<grid, 256>>>(buf, n_words, store_off, 1, sink);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0);
    k<U, MODE><<<grid, 256>>>(buf, n_words, store_off, 2 + r, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  const unsigned long long W = 4ull << 30;  // bytes walked per launch (>> the 126 MB L2)
  uint4 *buf;
  uint32_t *sink;
  CK(cudaMalloc(&buf, 2 * W));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(buf, 1, 2 * W));
  const unsigned long long n = W / 16;
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("{\"bytes_per_launch\":%llu,\"traffic\":\"2 x bytes (one load + one store per 16 B word)\",\"sms\":%d", W, sms);
  for (int per_sm : {2, 5, 8}) {
    const int grid = sms * per_sm;
    auto gbs = [&](float ms) { return 2.0 * (double)W / (ms * 1e-3) / 1e9; };
    printf(",\"ctas_per_sm_%d\":{", per_sm);
    printf("\"A_inplace_u4\":%.0f", gbs(run<4, 0>(buf, n, 0, grid, sink, 5)));
    printf(",\"B_copy_u4\":%.0f", gbs(run<4, 0>(buf, n, n, grid, sink, 5)));
    printf(",\"C_inplace_lag1_u4\":%.0f", gbs(run<4, 2>(buf, n, 0, grid, sink, 5)));
    printf(",\"E_inplace_nofeed_u4\":%.0f", gbs(run<4, 4>(buf, n, 0, grid, sink, 5)));
    printf(",\"F_inplace_u1\":%.0f", gbs(run<1, 0>(buf, n, 0, grid, sink, 5)));
    printf(",\"F_inplace_u8\":%.0f", gbs(run<8, 0>(buf, n, 0, grid, sink, 5)));
    printf(",\"C_inplace_lag1_u8\":%.0f", gbs(run<8, 2>(buf, n, 0, grid, sink, 5)));
    printf(",\"B_copy_u8\":%.0f}", gbs(run<8, 0>(buf, n, n, grid, sink, 5)));
  }
  printf(",\"unit\":\"GB/s of algorithmic traffic\"}\n");
  return 0;
}
