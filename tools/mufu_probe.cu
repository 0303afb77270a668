// MUFU.EX2 throughput on one SM partition mix: pure ex2 chains, and the Poisson sweep's mix
// (2 FFMA + EX2 + FADD + FFMA per element and particle), versus warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_dbg/mufu_probe tools/mufu_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float ex2(float t) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
  return r;
}

template <int MODE>
__global__ void probe(float* out, int iters, float a, float b) {
  float x[8], r0[2] = {0.f, 0.f}, r1[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = a * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {
        x[i] = ex2(x[i]);
      } else {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const float rate = ex2(fmaf(b, x[i], fmaf(a, x[i], -3.0f + q)));
          r0[q] += rate;
          r1[q] = fmaf(rate, x[i], r1[q]);
        }
        x[i] += 1e-6f;     // keeps the exponentials loop-variant
      }
    }
  }
  float s = r0[0] + r0[1] + r1[0] + r1[1];
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 12345.678f) out[0] = s;
}

int main() {
  float* out;
  cudaMalloc(&out, 4);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount, iters = 20000;
  for (int mode = 0; mode < 2; ++mode)
    for (int warps = 4; warps <= 32; warps *= 2) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) probe<0><<<sms, warps * 32>>>(out, iters, 1e-3f, 1e-4f);
        else probe<1><<<sms, warps * 32>>>(out, iters, 1e-3f, 1e-4f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double ex2s = (double)sms * warps * 32 * iters * 8 * (mode == 0 ? 1 : 2);
      printf("mode %d warps/SM %2d: %.3f ms, %.2f ex2 / ns / SM (%.1f per clk at 1.9 GHz)\n", mode, warps, ms,
             ex2s / (ms * 1e6) / sms, ex2s / (ms * 1e6) / sms / 1.9);
    }
  return 0;
}
