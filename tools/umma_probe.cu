// Standalone probe: which shared-memory layouts does tcgen05.mma kind::tf32 accept for K-major
// and MN-major operands on sm_100a? One CTA builds A [M x K] and B [N x K] images in smem with a
// given address function, issues K/8 MMAs and dumps the accumulator. Developer tool, not product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/_dbg/umma_probe tools/umma_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>

enum Major { KMAJ = 0, MNMAJ = 1 };
enum Lay { NONE = 0, SW128_32B = 1, SW128 = 2, SW64 = 4, SW32 = 6 };

struct Operand {
  int major, layout;
  uint32_t lbo, sbo;      // bytes
  uint32_t kstep;         // start-address advance per MMA (K = 8) in bytes
};

// byte offset of element (mn, k) inside the operand image
__host__ __device__ inline uint32_t elem_offset(const Operand& o, int mn, int k) {
  uint32_t off;
  if (o.major == KMAJ) {
    if (o.layout == NONE) {
      // ((8,m),(T,2k)) : core matrix 8 rows x 16 B contiguous; SBO between 8-row groups, LBO between 16-B k chunks
      off = (mn / 8) * o.sbo + (k / 4) * o.lbo + (mn % 8) * 16 + (k % 4) * 4;
    } else if (o.layout == SW128_32B) {  // hypothetical: K-major read of a 32B-atom swizzled image
      off = (k / 32) * o.lbo + (mn / 8) * o.sbo + (mn % 8) * 128 + (k % 32) * 4;
      off ^= ((off >> 7) & 3u) << 5;
    } else {  // SW128: rows of 128 B (32 fp32 of K), 8-row groups SBO apart; k < 32 per atom, atoms LBO apart
      off = (k / 32) * o.lbo + (mn / 8) * o.sbo + (mn % 8) * 128 + (k % 32) * 4;
      off ^= ((off >> 7) & 7u) << 4;
    }
  } else {
    if (o.layout == NONE) {
      // ((T,1,m),(8,k)) : core matrix 8 k-rows x 16 B (4 mn) contiguous; SBO between mn chunks of 4, LBO between k groups of 8
      off = (mn / 4) * o.sbo + (k / 8) * o.lbo + (k % 8) * 16 + (mn % 4) * 4;
    } else if (o.layout == SW128) {
      off = (mn / 32) * o.lbo + (k / 8) * o.sbo + (k % 8) * 128 + (mn % 32) * 4;
      off ^= ((off >> 7) & 7u) << 4;
    } else {  // SW128_32B: 4 k-rows of 128 B per group, Swizzle<2,5,2>
      off = (mn / 32) * o.lbo + (k / 4) * o.sbo + (k % 4) * 128 + (mn % 32) * 4;
      off ^= ((off >> 7) & 3u) << 5;
    }
  }
  return off;
}

__device__ inline uint64_t make_desc(uint32_t addr, const Operand& o) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((o.lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((o.sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(o.layout & 7) << 61;
  return d;
}

__global__ void probe(const float* A, const float* B, int M, int N, int K, Operand oa, Operand ob,
                      float* D /* [128 lanes][N] */) {
  extern __shared__ uint8_t raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = ((uint32_t)__cvta_generic_to_shared(raw) + 1023u) & ~1023u;
  uint8_t* g = raw + (base - (uint32_t)__cvta_generic_to_shared(raw));
  const uint32_t offA = 0, offB = 64 * 1024;
  for (int i = threadIdx.x; i < 128 * 1024; i += blockDim.x) g[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
    const int mn = i / K, k = i % K;
    *reinterpret_cast<float*>(g + offA + elem_offset(oa, mn, k)) = A[i];
  }
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int mn = i / K, k = i % K;
    *reinterpret_cast<float*>(g + offB + elem_offset(ob, mn, k)) = B[i];
  }
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(
        (uint32_t)__cvta_generic_to_shared(&tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  // zero the accumulator region first so "no write" is distinguishable from "wrote zeros"
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)oa.major << 15) |
                           ((uint32_t)ob.major << 16) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)(M >> 4) << 24);
    for (int ks = 0; ks < K / 8; ++ks) {
      const uint64_t da = make_desc(base + offA + ks * oa.kstep, oa);
      const uint64_t db = make_desc(base + offB + ks * ob.kstep, ob);
      const uint32_t accum = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(accum)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr)
                 : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DN;\n\tbra W;\n\tDN:\n\t}" ::"r"(bar_addr), "r"(0)
      : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x / 32;
  for (int c = 0; c < N; c += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]),
                   "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) D[threadIdx.x * N + c + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

static float tf32(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u = (u + 0x1000u) & 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

static const char* lname(int l) {
  switch (l) { case NONE: return "NONE"; case SW128: return "SW128"; case SW128_32B: return "SW128_32B"; }
  return "?";
}

int run(int M, int N, int K, Operand oa, Operand ob) {
  std::vector<float> A(M * K), B(N * K), Dref(M * N);
  srand(1);
  for (auto& v : A) v = tf32((float)rand() / RAND_MAX - 0.5f);
  for (auto& v : B) v = tf32((float)rand() / RAND_MAX - 0.5f);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double t = 0;
      for (int k = 0; k < K; ++k) t += (double)A[m * K + k] * B[n * K + k];
      Dref[m * N + n] = (float)t;
    }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, 128 * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, 128 * N * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 130 * 1024);
  probe<<<1, 128, 130 * 1024>>>(dA, dB, M, N, K, oa, ob, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  CUDA error: %s\n", cudaGetErrorString(e));
    return 1;
  }
  std::vector<float> D(128 * N);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  // accumulator row m lives on lane m (M=128) or (m%16) + 32*(m/16) (M=64)
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m) {
    const int lane = M == 128 ? m : (m % 16) + 32 * (m / 16);
    for (int n = 0; n < N; ++n) {
      maxerr = fmax(maxerr, fabs((double)D[lane * N + n] - Dref[m * N + n]));
      maxref = fmax(maxref, fabs((double)Dref[m * N + n]));
    }
  }
  double absmax = 0;
  for (float v : D) absmax = fmax(absmax, fabs((double)v));
  printf("M=%d N=%d K=%d  A:%s/%s lbo=%u sbo=%u kstep=%u  B:%s/%s lbo=%u sbo=%u kstep=%u  -> max err %.3e (ref max %.3e, out max %.3e) %s\n",
         M, N, K, oa.major ? "MN" : "K", lname(oa.layout), oa.lbo, oa.sbo, oa.kstep,
         ob.major ? "MN" : "K", lname(ob.layout), ob.lbo, ob.sbo, ob.kstep, maxerr, maxref, absmax,
         maxerr < 1e-4 * fmax(maxref, 1e-30) + 1e-5 ? "OK" : "WRONG");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return 0;
}

int main() {
  // reference points that must work: K-major SW128 both (the eta GEMM of the kernel)
  Operand k128 = {KMAJ, SW128, 16, 1024, 32};
  // K-major, no swizzle: 8x16B core matrices; k chunks (16 B) LBO apart, 8-row groups SBO apart.
  // image A: [row group][k chunk][8][16B]: SBO = (K/4)*128, LBO = 128 ; kstep = 2 chunks = 2*LBO
  auto knone = [](int K) { Operand o = {KMAJ, NONE, 128, (uint32_t)(K / 4) * 128, 256}; return o; };
  // MN-major, standard SW128 (what the first kernel used): 32 mn contiguous, 8 k rows per 1024 B
  auto mn128 = [](int K) { Operand o = {MNMAJ, SW128, (uint32_t)(K / 8) * 1024, 1024, 1024}; return o; };
  // MN-major, SW128 with 32B atomicity: 4 k rows per 512-B group
  auto mn128_32 = [](int K) { Operand o = {MNMAJ, SW128_32B, (uint32_t)(K / 4) * 512, 512, 1024}; return o; };
  // MN-major, no swizzle: core matrix 8 k x 16 B; mn chunks (4 elems) SBO apart, k groups LBO apart
  // image: [k group][mn chunk][8][16B]: SBO = 128, LBO = (MN/4)*128 ; kstep = LBO
  auto mnnone = [](int MN) { Operand o = {MNMAJ, NONE, (uint32_t)(MN / 4) * 128, 128, (uint32_t)(MN / 4) * 128}; return o; };
  // the same unswizzled image read as K-major by the *other* GEMM: element (row r, feature f) at
  // (r/8)*G + (f/4)*128 + (r%8)*16 + (f%4)*4 with G = (F/4)*128  == mnnone(F) with mn=f,k=r
  {
    printf("--- K-major read with layout type SW128_32B (same image as MN-major SW128_32B of the transpose)\n");
    Operand k32b = {KMAJ, SW128_32B, 16, 1024, 32};
    run(128, 64, 32, k32b, k128);
    run(128, 64, 8, k32b, k128);
    Operand k32b2 = {KMAJ, SW128_32B, 16, 512, 32};
    run(128, 64, 8, k32b2, k128);
  }
  printf("--- sanity: K-major SW128 x K-major SW128\n");
  run(128, 64, 32, k128, k128);
  run(64, 64, 32, k128, k128);
  printf("--- K-major NONE\n");
  run(128, 64, 32, knone(32), knone(32));
  run(128, 64, 64, knone(64), knone(64));
  printf("--- MN-major variants for A, B K-major SW128\n");
  run(64, 64, 16, mn128(16), k128);
  run(64, 64, 16, mn128_32(16), k128);
  run(64, 64, 16, mnnone(64), k128);
  printf("--- MN-major variants for both\n");
  run(64, 64, 16, mn128(16), mn128(16));
  run(64, 64, 16, mn128_32(16), mn128_32(16));
  run(64, 64, 16, mnnone(64), mnnone(64));
  run(64, 64, 128, mnnone(64), mnnone(64));
  run(128, 64, 128, mnnone(128), mnnone(64));
  printf("--- mixed: A MN NONE, B MN SW128_32B and vice versa\n");
  run(64, 64, 16, mnnone(64), mn128_32(16));
  run(64, 64, 16, mn128_32(16), mnnone(64));
  return 0;
}
