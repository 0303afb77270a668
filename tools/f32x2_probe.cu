#include <cstdio>
#include <cstdint>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  uint64_t ra = *reinterpret_cast<uint64_t*>(&a), rb = *reinterpret_cast<uint64_t*>(&b), rc = *reinterpret_cast<uint64_t*>(&c), rd;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  uint64_t ra = *reinterpret_cast<uint64_t*>(&a), rb = *reinterpret_cast<uint64_t*>(&b), rd;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  return *reinterpret_cast<float2*>(&rd);
}
template <int MODE>
__global__ void k(float* out, int iters, float seed) {
  float2 a[8], b[8];
  for (int i = 0; i < 8; ++i) { a[i] = make_float2(seed + i, seed - i); b[i] = make_float2(0.f, 0.f); }
  float2 m = make_float2(0.999f, 1.001f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, b[i].x); a[i].y = fmaf(a[i].y, m.y, b[i].y); b[i].x += a[i].x; b[i].y += a[i].y; }
      else { a[i] = fma2(a[i], m, b[i]); b[i] = add2(b[i], a[i]); }
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y + b[i].x + b[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 2; ++rep) {
    cudaEventRecord(e0);
    if (mode == 0) k<0><<<148 * 8, 256>>>(out, 20000, 1.0f); else k<1><<<148 * 8, 256>>>(out, 20000, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 148.0 * 8 * 256 * 20000.0 * 8 * 2 * 3;   // per i: 2 fma (4 flop) + 2 add (2 flop)
    printf("mode %d: %.3f ms, %.1f TFLOP/s fp32 (fma=2)\n", mode, ms, flops / ms / 1e9);
  }
  return 0;
}
