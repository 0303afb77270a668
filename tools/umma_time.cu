// Standalone timing probe: cycles per tcgen05.mma kind::tf32 for operand configurations.
// One CTA, one issuing thread, NREP back-to-back MMAs into the same accumulator, clock64 from
// first issue to commit arrival. Developer tool.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

struct Cfg { int M, N, a_major, a_layout, b_major, b_layout, a_tmem, nrep, nacc; uint32_t a_lbo, a_sbo, b_lbo, b_sbo; };

__device__ inline uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, int layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred;
}

__global__ void timek(Cfg c, long long* out) {
  extern __shared__ uint8_t raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = ((uint32_t)__cvta_generic_to_shared(raw) + 1023u) & ~1023u;
  uint8_t* g = raw + (base - (uint32_t)__cvta_generic_to_shared(raw));
  for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(g)[i] = 1.0f;
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32 && elect_one()) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)c.a_major << 15) |
                           ((uint32_t)c.b_major << 16) | ((uint32_t)(c.N >> 3) << 17) |
                           ((uint32_t)(c.M >> 4) << 24);
    const uint64_t da0 = make_desc(base, c.a_lbo, c.a_sbo, c.a_layout);
    const uint64_t db0 = make_desc(base + 80 * 1024, c.b_lbo, c.b_sbo, c.b_layout);
    const uint32_t a_lo = (uint32_t)da0, a_hi = (uint32_t)(da0 >> 32);
    const uint32_t b_lo = (uint32_t)db0, b_hi = (uint32_t)(db0 >> 32);
    const uint32_t dstep = c.nacc > 1 ? (uint32_t)c.N : 0u;
    const long long t0 = clock64();
    for (int r = 0; r < c.nrep; r += 16) {
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        // 16 operand slices 1 KB apart (64 in descriptor units), accumulators round-robin over 4
        const uint32_t d = tmem + (uint32_t)(u & 3) * dstep;
        if (c.a_tmem) {
          asm volatile(
              "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2,%3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}" ::"r"(d),
              "r"(tmem + 256 + (uint32_t)u * 8), "r"(b_lo + u * 64), "r"(b_hi), "r"(idesc), "r"(1)
              : "memory");
        } else {
          asm volatile(
              "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1,%2};\n\tmov.b64 db, {%3,%4};\n\t"
              "setp.ne.b32 p, %6, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(d),
              "r"(a_lo + u * 64), "r"(a_hi), "r"(b_lo + u * 64), "r"(b_hi), "r"(idesc), "r"(1)
              : "memory");
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr)
                 : "memory");
    asm volatile(
        "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DN;\n\tbra W;\n\tDN:\n\t}" ::"r"(bar_addr), "r"(0)
        : "memory");
    out[0] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

static void run(const char* name, Cfg c) {
  long long* d; cudaMalloc(&d, 8);
  cudaFuncSetAttribute(timek, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long best = 1LL << 60;
  for (int rep = 0; rep < 3; ++rep) {
    timek<<<1, 128, 200 * 1024>>>(c, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s CUDA error %s\n", name, cudaGetErrorString(e)); return; }
    long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    if (h < best) best = h;
  }
  const double per = (double)best / c.nrep;
  const double macs = (double)c.M * c.N * 8;
  printf("%-44s M=%3d N=%3d : %7.1f cycles/MMA  (%6.0f MAC/cycle)\n", name, c.M, c.N, per, macs / per);
  cudaFree(d);
}

int main() {
  const int R = 1024;
  for (int nacc = 1; nacc <= 4; nacc *= 4) {
    printf("--- %s\n", nacc == 1 ? "one accumulator (dependent chain)" : "4 accumulators round-robin");
    run("A K/SW128   B K/SW128", {128, 64, 0, 2, 0, 2, 0, R, nacc, 16, 1024, 16, 1024});
    run("A K/SW128   B K/SW128", {64, 64, 0, 2, 0, 2, 0, R, nacc, 16, 1024, 16, 1024});
    run("A K/SW128   B K/SW128", {128, 128, 0, 2, 0, 2, 0, R, nacc, 16, 1024, 16, 1024});
    run("A K/SW128   B K/SW128", {64, 128, 0, 2, 0, 2, 0, R, nacc, 16, 1024, 16, 1024});
    if (nacc == 1) run("A K/SW128   B K/SW128", {128, 256, 0, 2, 0, 2, 0, R, nacc, 16, 1024, 16, 1024});
    if (nacc == 1) run("A K/SW128   B K/SW128", {64, 256, 0, 2, 0, 2, 0, R, nacc, 16, 1024, 16, 1024});
    run("A MN/SW32B  B MN/SW32B", {64, 64, 1, 1, 1, 1, 0, R, nacc, 16384, 512, 16384, 512});
    run("A MN/SW32B  B MN/SW32B", {128, 64, 1, 1, 1, 1, 0, R, nacc, 16384, 512, 16384, 512});
    run("A MN/SW32B  B K/SW128", {64, 64, 1, 1, 0, 2, 0, R, nacc, 16384, 512, 16, 1024});
    run("A K/SW128   B MN/SW32B", {128, 64, 0, 2, 1, 1, 0, R, nacc, 16, 1024, 16384, 512});
    run("A K/SW128   B MN/SW32B", {128, 128, 0, 2, 1, 1, 0, R, nacc, 16, 1024, 16384, 512});
    run("A TMEM      B K/SW128", {64, 64, 0, 2, 0, 2, 1, R, nacc, 16, 1024, 16, 1024});
    run("A TMEM      B K/SW128", {128, 64, 0, 2, 0, 2, 1, R, nacc, 16, 1024, 16, 1024});
    run("A TMEM      B K/SW128", {64, 128, 0, 2, 0, 2, 1, R, nacc, 16, 1024, 16, 1024});
    run("A TMEM      B K/SW128", {128, 128, 0, 2, 0, 2, 1, R, nacc, 16, 1024, 16, 1024});   // hi/lo particle slots
    run("A TMEM      B MN/SW32B", {64, 64, 0, 2, 1, 1, 1, R, nacc, 16, 1024, 16384, 512});
    run("A TMEM      B MN/SW32B", {128, 64, 0, 2, 1, 1, 1, R, nacc, 16, 1024, 16384, 512});
    run("A TMEM      B MN/SW32B", {64, 128, 0, 2, 1, 1, 1, R, nacc, 16, 1024, 16384, 512});
  }
  // small-N shapes of the rows-on-lanes kernel (dense_tcr.cuh): particles on N
  for (int nacc = 1; nacc <= 4; nacc *= 4) {
    printf("--- small N, %s\n", nacc == 1 ? "one accumulator" : "4 accumulators round-robin");
    for (int N = 16; N <= 64; N *= 2) {
      run("A K/SW128   B K/SW128", {128, N, 0, 2, 0, 2, 0, R, nacc, 16, 1024, 16, 1024});
      run("A K/SW128   B K/SW128", {64, N, 0, 2, 0, 2, 0, R, nacc, 16, 1024, 16, 1024});
      run("A MN/SW32B  B K/SW128", {64, N, 1, 1, 0, 2, 0, R, nacc, 16384, 512, 16, 1024});
      run("A MN/SW32B  B K/SW128", {128, N, 1, 1, 0, 2, 0, R, nacc, 16384, 512, 16, 1024});
      run("A TMEM      B K/SW128", {128, N, 0, 2, 0, 2, 1, R, nacc, 16, 1024, 16, 1024});
      run("A TMEM      B K/SW128", {64, N, 0, 2, 0, 2, 1, R, nacc, 16, 1024, 16, 1024});
    }
  }
  return 0;
}
