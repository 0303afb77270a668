// Probe TMA (cp.async.bulk.tensor.2d) on B200: (1) does tensor-map data type TFLOAT32 round fp32
// to tf32 on the way into shared memory, and how; (2) do SWIZZLE_128B and SWIZZLE_128B_ATOM_32B
// produce exactly the shared-memory images the tcgen05 descriptors of dense_tc.cuh expect.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tm, uint32_t* out, int row0) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t base = ((uint32_t)__cvta_generic_to_shared(raw) + 1023u) & ~1023u;
  uint8_t* g = raw + (base - (uint32_t)__cvta_generic_to_shared(raw));
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(2 * 16384) : "memory");
    for (int a = 0; a < 2; ++a)   // two boxes: features [0,32) and [32,64)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                   ::"r"(base + a * 16384), "l"(&tm), "r"(a * 32), "r"(row0), "r"(bar_addr) : "memory");
  }
  asm volatile(
      "{\n\t.reg .pred P1;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DN;\n\tbra W;\n\tDN:\n\t}" ::"r"(bar_addr), "r"(0) : "memory");
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) out[i] = reinterpret_cast<uint32_t*>(g)[i];
}

static uint32_t bits(float x) { uint32_t u; memcpy(&u, &x, 4); return u; }

int main() {
  EncodeFn encode; cudaDriverEntryPointQueryResult qres;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres);
  const int N = 300, P = 64;   // 300 rows: the third tile is ragged (44 rows)
  std::vector<float> X(N * P);
  srand(3);
  for (auto& v : X) { uint32_t u = (uint32_t)rand() ^ ((uint32_t)rand() << 16); u = (u & 0x007FFFFFu) | 0x3F800000u; memcpy(&v, &u, 4); v = (rand() & 1) ? v : -v; }
  // a few exact ties: low 13 bits = 0x1000
  for (int i = 0; i < 16; ++i) { uint32_t u = bits(X[i]); u = (u & 0xFFFFE000u) | 0x1000u; memcpy(&X[i], &u, 4); }
  float* dX; cudaMalloc(&dX, X.size() * 4); cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  uint32_t* dOut; cudaMalloc(&dOut, 8192 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  for (int mode = 0; mode < 4; ++mode) {
    const CUtensorMapDataType dt = (mode & 1) ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    const CUtensorMapSwizzle sw = (mode & 2) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)P, (cuuint64_t)N};
    cuuint64_t strides[1] = {(cuuint64_t)P * 4};
    cuuint32_t box[2] = {32, 128};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tm, dt, 2, dX, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("mode %d encode failed %d\n", mode, (int)r); continue; }
    for (int tile = 0; tile < 3; tile += 2) {
      cudaMemset(dOut, 0xFF, 8192 * 4);
      probe<<<1, 128, 40 * 1024>>>(tm, dOut, tile * 128);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d CUDA error %s\n", mode, cudaGetErrorString(e)); return 1; }
      std::vector<uint32_t> out(8192);
      cudaMemcpy(out.data(), dOut, 8192 * 4, cudaMemcpyDeviceToHost);
      long n_exact = 0, n_rna = 0, n_rne = 0, n_trunc = 0, n_other = 0, n_zero_oob = 0, n = 0, layout_bad = 0;
      for (int r_ = 0; r_ < 128; ++r_) for (int f = 0; f < 64; ++f) {
        uint32_t off = (f / 32) * 16384 + r_ * 128 + (f % 32) * 4;
        if (mode & 2) off ^= ((off >> 7) & 3u) << 5; else off ^= ((off >> 7) & 7u) << 4;
        const uint32_t got = out[off / 4];
        const int row = tile * 128 + r_;
        if (row >= N) { n_zero_oob += (got == 0); continue; }
        const uint32_t src = bits(X[row * P + f]);
        const uint32_t rna = (src + 0x1000u) & 0xFFFFE000u;
        const uint32_t rne = (src + 0x0FFFu + ((src >> 13) & 1u)) & 0xFFFFE000u;
        const uint32_t tr = src & 0xFFFFE000u;
        ++n;
        if (got == src) ++n_exact;
        if (got == rna) ++n_rna;
        if (got == rne) ++n_rne;
        if (got == tr) ++n_trunc;
        if (got != src && got != rna && got != rne && got != tr) { ++n_other; ++layout_bad; }
      }
      printf("dtype %-8s swizzle %-14s tile %d: n=%ld exact=%ld rna=%ld rne=%ld trunc=%ld other=%ld oob_zero=%ld\n",
             (mode & 1) ? "TFLOAT32" : "FLOAT32", (mode & 2) ? "128B_ATOM_32B" : "128B", tile, n, n_exact, n_rna, n_rne, n_trunc, n_other, n_zero_oob);
    }
  }
  return 0;
}
