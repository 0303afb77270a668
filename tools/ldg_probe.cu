// How many producer warps x outstanding LDG.128 per thread does one SM need to saturate HBM?
// Persistent grid of 148 CTAs, W warps each; every thread issues U independent 16-byte loads,
// then consumes them (xor-reduce), looping over a 12.8 GB buffer.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
template <int U>
__global__ void stream(const float4* __restrict__ src, size_t n_vec, float* sink) {
  const size_t per_iter = (size_t)gridDim.x * blockDim.x * U;
  float acc = 0.f;
  for (size_t base = (size_t)blockIdx.x * blockDim.x * U; base + (size_t)blockDim.x * U <= n_vec; base += per_iter) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const float4* p = src + base + (size_t)u * blockDim.x + threadIdx.x;
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(p));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) sink[0] = acc;
}
template <int U>
void run(const float4* d, size_t n_vec, float* sink, int warps, int reps = 3) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  stream<U><<<148, warps * 32>>>(d, n_vec, sink);
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) stream<U><<<148, warps * 32>>>(d, n_vec, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  printf("warps/SM %2d  x  %2d LDG.128/thread (%3d KB requested per SM per round): %7.1f GB/s\n", warps, U, warps * 32 * U * 16 / 1024, n_vec * 16.0 / ms / 1e6);
}
int main() {
  const size_t bytes = 12800ull << 20; const size_t n_vec = bytes / 16;
  float4* d; cudaMalloc(&d, bytes); cudaMemset(d, 0, bytes);
  float* sink; cudaMalloc(&sink, 4);
  run<16>(d, n_vec, sink, 8); run<8>(d, n_vec, sink, 8); run<4>(d, n_vec, sink, 8);
  run<16>(d, n_vec, sink, 4); run<32>(d, n_vec, sink, 4);
  run<8>(d, n_vec, sink, 16); run<4>(d, n_vec, sink, 16); run<16>(d, n_vec, sink, 16);
  run<4>(d, n_vec, sink, 32); run<8>(d, n_vec, sink, 32); run<2>(d, n_vec, sink, 32);
  run<16>(d, n_vec, sink, 12); run<8>(d, n_vec, sink, 12); run<12>(d, n_vec, sink, 12);
  // the same loads against an L2-resident buffer (64 MB of the 126 MB L2): what register-staged
  // loads could sustain behind an L2 prefetch
  printf("--- 64 MB buffer (L2-resident)\n");
  const size_t small = (64ull << 20) / 16;
  run<4>(d, small, sink, 8, 200); run<8>(d, small, sink, 8, 200); run<2>(d, small, sink, 16, 200);
  run<4>(d, small, sink, 16, 200); run<8>(d, small, sink, 16, 200); run<4>(d, small, sink, 32, 200);
  return 0;
}
