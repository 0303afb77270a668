// Probe: tcgen05.mma kind::tf32 with the A operand in tensor memory (M = 64), B in shared memory
// K-major (SW128) or MN-major (SW128_32B). Checks the TMEM layout of A: element A[m][k] is
// written by tcgen05.st 32x32b to lane (m%16)+32*(m/16), column col0 + k.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <vector>

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
  asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xFFFFFFFF;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred;
}

// A [64][K], B [N][K]; b_mn: 0 -> K-major SW128 image, 1 -> MN-major SW128_32B image
__global__ void probe(const float* A, const float* B, int N, int K, int b_mn, int m_is_128, float* D) {
  extern __shared__ uint8_t raw[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = ((uint32_t)__cvta_generic_to_shared(raw) + 1023u) & ~1023u;
  uint8_t* g = raw + (base - (uint32_t)__cvta_generic_to_shared(raw));
  const int M = m_is_128 ? 128 : 64;
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    uint32_t off;
    if (!b_mn) {  // K-major SW128: atoms of 32 k, atom stride N*128
      off = (k / 32) * (N * 128) + n * 128 + (k % 32) * 4;
      off ^= ((off >> 7) & 7u) << 4;
    } else {      // MN-major SW128_32B: atoms of 32 n, atom stride K*128
      off = (n / 32) * (K * 128) + k * 128 + (n % 32) * 4;
      off ^= ((off >> 7) & 3u) << 5;
    }
    *reinterpret_cast<float*>(g + off) = B[i];
  }
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
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  // write A into TMEM columns [256, 256+K): row m on lane (m%16)+32*(m/16) for M=64, lane m for M=128
  const uint32_t a_col = 256;
  {
    int m = -1;
    if (M == 128) m = threadIdx.x;
    else if (lane < 16) m = warp * 16 + lane;
    for (int k0 = 0; k0 < K; k0 += 8) {
      uint32_t v[8];
      for (int i = 0; i < 8; ++i) v[i] = m >= 0 ? __float_as_uint(A[m * K + k0 + i]) : 0u;
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(
                       tmem + ((uint32_t)(warp * 32) << 16) + a_col + k0),
                   "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                   : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (0u << 15) | ((uint32_t)b_mn << 16) |
                           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int ks = 0; ks < K / 8; ++ks) {
      uint64_t db;
      if (!b_mn) db = make_desc(base + (ks / 4) * (N * 128) + (ks % 4) * 32, 16, 1024, 2);
      else db = make_desc(base + ks * 1024, K * 128, 512, 1);
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
          "r"(tmem + a_col + ks * 8), "l"(db), "r"(idesc), "r"((uint32_t)(ks > 0))
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DN;\n\tbra W;\n\tDN:\n\t}" ::"r"(bar_addr), "r"(0) : "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c = 0; c < N; c += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 8; ++i) D[threadIdx.x * N + c + i] = __uint_as_float(v[i]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

static float tf32(float x) { uint32_t u; memcpy(&u, &x, 4); u = (u + 0x1000u) & 0xFFFFE000u; memcpy(&x, &u, 4); return x; }

static void run(int M, int N, int K, int b_mn) {
  std::vector<float> A(M * K), B(N * K), Dref(M * N);
  srand(2);
  for (auto& v : A) v = tf32((float)rand() / RAND_MAX - 0.5f);
  for (auto& v : B) v = tf32((float)rand() / RAND_MAX - 0.5f);
  for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
    double t = 0; for (int k = 0; k < K; ++k) t += (double)A[m * K + k] * B[n * K + k];
    Dref[m * N + n] = (float)t;
  }
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, 128 * N * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, 128 * N * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 140 * 1024);
  probe<<<1, 128, 140 * 1024>>>(dA, dB, N, K, b_mn, M == 128, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("M=%d N=%d K=%d b_mn=%d CUDA error: %s\n", M, N, K, b_mn, cudaGetErrorString(e)); return; }
  std::vector<float> D(128 * N);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m) {
    const int lane = M == 128 ? m : (m % 16) + 32 * (m / 16);
    for (int n = 0; n < N; ++n) {
      maxerr = fmax(maxerr, fabs((double)D[lane * N + n] - Dref[m * N + n]));
      maxref = fmax(maxref, fabs((double)Dref[m * N + n]));
    }
  }
  printf("TS M=%d N=%d K=%d B %s : max err %.3e (ref max %.3e) %s\n", M, N, K, b_mn ? "MN/SW128_32B" : "K/SW128",
         maxerr, maxref, maxerr < 1e-4 * maxref + 1e-6 ? "OK" : "WRONG");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

int main() {
  run(64, 128, 64, 0);   // eta^T = Theta . X^T : A = Theta [64 x 64] in TMEM, B = X tile K-major [128 rows x 64]
  run(64, 64, 128, 1);   // G^T = R^T . X     : A = R^T [64 x 128 rows] in TMEM, B = X MN-major [64 feats x 128 rows]
  run(128, 64, 64, 0);
  run(128, 64, 128, 1);
  run(64, 256, 64, 0);
  return 0;
}
