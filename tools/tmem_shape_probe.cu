// Probe the register mapping of tcgen05.ld/st .16x256b: write lane*1000+col with 32x32b stores,
// read back with 16x256b.x4 and print which (lane, col) each thread register received.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
__global__ void probe(float* out /* [128 threads][16] */, float* out2 /* [128][8] readback via 32x32b after 16x256b store */) {
  __shared__ uint32_t tmem_slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c0 = 0; c0 < 32; c0 += 8) {
    uint32_t v[8];
    for (int i = 0; i < 8; ++i) v[i] = __float_as_uint((float)(threadIdx.x * 1000 + c0 + i));
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem + lane_base + c0),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncwarp();
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(tmem + lane_base));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; ++i) out[threadIdx.x * 16 + i] = __uint_as_float(r[i]);
  // now store r + 0.5 back with the same shape and read via 32x32b to verify st mapping symmetry
  for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) + 0.5f);
  asm volatile("tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(tmem + lane_base),
               "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
               "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncwarp();
  uint32_t w[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(tmem + lane_base));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 8; ++i) out2[threadIdx.x * 8 + i] = __uint_as_float(w[i]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}
// 16x256b.x4 load addressed at lane 16 of each warp's quadrant
__global__ void probe_hi(float* out /* [128 threads][16] */) {
  __shared__ uint32_t tmem_slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int warp = threadIdx.x / 32;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  for (int c0 = 0; c0 < 32; c0 += 8) {
    uint32_t v[8];
    for (int i = 0; i < 8; ++i) v[i] = __float_as_uint((float)(threadIdx.x * 1000 + c0 + i));
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem + lane_base + c0),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncwarp();
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(tmem + lane_base + (16u << 16)));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; ++i) out[threadIdx.x * 16 + i] = __uint_as_float(r[i]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}
// 16x128b.x4 store: which (lane, column) does register i of thread t land in? Read back via 32x32b.
__global__ void probe_st128(float* out /* [128 threads][16 cols] */) {
  __shared__ uint32_t tmem_slot;
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  uint32_t z[16];
  for (int i = 0; i < 16; ++i) z[i] = __float_as_uint(-1.0f);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(tmem + lane_base),
               "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]),
               "r"(z[8]), "r"(z[9]), "r"(z[10]), "r"(z[11]), "r"(z[12]), "r"(z[13]), "r"(z[14]), "r"(z[15]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncwarp();
  uint32_t v[8];
  for (int i = 0; i < 8; ++i) v[i] = __float_as_uint((float)(lane * 100 + i));     // thread-in-warp * 100 + register
  asm volatile("tcgen05.st.sync.aligned.16x128b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem + lane_base),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncwarp();
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                 "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(tmem + lane_base));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int i = 0; i < 16; ++i) out[threadIdx.x * 16 + i] = __uint_as_float(r[i]);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}
int main() {
  float *d, *d2; cudaMalloc(&d, 128 * 16 * 4); cudaMalloc(&d2, 128 * 8 * 4);
  probe<<<1, 128>>>(d, d2);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  float h[128 * 16], h2[128 * 8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(h2, d2, sizeof(h2), cudaMemcpyDeviceToHost);
  int ok = 1;
  for (int t = 32; t < 64; ++t) {   // warp 1
    if (t < 40 || t == 63) { printf("thread %3d:", t); for (int i = 0; i < 16; ++i) printf(" %6.0f", h[t * 16 + i]); printf("\n"); }
    const int l = t % 32;
    for (int i = 0; i < 16; ++i) {
      const int g = i / 4, j = i % 4;
      const int lane = 32 + l / 4 + (j >= 2 ? 8 : 0), col = 8 * g + 2 * (l % 4) + (j & 1);
      if (h[t * 16 + i] != (float)(lane * 1000 + col)) ok = 0;
    }
  }
  printf("assumed mapping (row = t/4 [+8], col = 8g + 2(t%%4) + {0,1}) %s\n", ok ? "CONFIRMED" : "WRONG");
  int ok2 = 1;
  for (int t = 0; t < 128; ++t) for (int i = 0; i < 8; ++i) {
    const float expect = (t % 32) < 16 ? t * 1000 + i + 0.5f : t * 1000 + i;
    if (h2[t * 8 + i] != expect) ok2 = 0;
  }
  printf("16x256b store round trip (lanes 0-15 updated, 16-31 untouched) %s\n", ok2 ? "CONFIRMED" : "WRONG");
  // second launch: the same shape addressed at lane 16 of the quadrant (upper half-quadrant)
  probe_hi<<<1, 128>>>(d);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("16x256b at lane offset 16: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int ok3 = 1;
  for (int t = 0; t < 128; ++t) {
    const int l = t % 32, q = t / 32;
    for (int i = 0; i < 16; ++i) {
      const int g = i / 4, j = i % 4;
      const int lane = 32 * q + 16 + l / 4 + (j >= 2 ? 8 : 0), col = 8 * g + 2 * (l % 4) + (j & 1);
      if (h[t * 16 + i] != (float)(lane * 1000 + col)) ok3 = 0;
    }
  }
  printf("thread  33:"); for (int i = 0; i < 16; ++i) printf(" %6.0f", h[33 * 16 + i]); printf("\n");
  printf("16x256b at lane offset 16 reads lanes 16-31 with the same register mapping %s\n", ok3 ? "CONFIRMED" : "WRONG");
  // third launch: 16x128b.x4 store mapping (dense_th.cuh writes packed fp16 score pairs with it)
  probe_st128<<<1, 128>>>(d);
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("16x128b store: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int ok4 = 1;
  for (int lane = 0; lane < 32; ++lane) {          // warp 0's quadrant: lane = TMEM lane, 16 columns
    if (lane < 2 || lane == 8 || lane == 16) { printf("lane %2d:", lane); for (int c = 0; c < 16; ++c) printf(" %5.0f", h[lane * 16 + c]); printf("\n"); }
    for (int c = 0; c < 16; ++c) {
      float expect = -1.0f;                          // lanes 16-31 untouched
      if (lane < 16) {
        const int t = 4 * (lane % 8) + c % 4, g = c / 4, j = lane / 8;   // assumed: reg 2g+j of thread t -> lane t/4 + 8j, col 4g + t%4
        expect = (float)(t * 100 + 2 * g + j);
      }
      if (h[lane * 16 + c] != expect) ok4 = 0;
    }
  }
  printf("16x128b.x4 store: register 2g+j of thread t -> lane t/4 + 8j, column 4g + t%%4 %s\n", ok4 ? "CONFIRMED" : "WRONG");
  return 0;
}
