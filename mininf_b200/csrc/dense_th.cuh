// Dense-link sweep, tcgen05 kind::f16 variant (MNF_DENSE_F16) for sm_100a: the black-box
// per-(particle, observation) kernel of config C2 (p = 64, S <= 64) at half the tensor-pipe and
// tensor-memory cost of dense_tc.cuh.
//
// Why a second operand format. A kind::tf32 MMA covers K = 8 per instruction, a kind::f16 MMA
// K = 16 at the same cycle count (tools/umma_time.cu; B300_MICROARCH.md: floor = max(M,128)*N/256
// cycles per instruction, K_per_mma = 32 bytes / element size), and every instruction makes one
// read-modify-write pass over its accumulator in tensor memory. dense_tc.cuh spends 24 MMAs =
// 1273 tensor cycles and ~1.1 MB of TMEM traffic per 128-row tile, which under the board's power
// cap (SM clock ~1.3-1.5 GHz) is MORE than the tile's HBM time (~1050 cycles): that kernel is
// tensor-bound, not HBM-bound. IEEE half precision has the SAME 11-bit significand as TF32, so
// with operands in fp16 the products are as exact as before, the tile costs 12 MMAs = 637 cycles,
// and - since for 16-bit operands the K-major SWIZZLE_128B image of X (operand of the eta
// product) and the MN-major SWIZZLE_128B image (operand of the gradient product) are the same
// bytes - ONE 16 KB image per tile serves both products: X leaves HBM once and is never re-read
// through L2.
//
//   fp32 tile [128 x 64]  --TMA-->  staging ring (SWIZZLE_128B, 32 KB)
//   staging  --eight conversion warps: LDS.128 -> cvt.rn.f16x2.f32 -> STS.128-->  fp16 image (16 KB)
//   eta^T[2S x 128] = Theta[2S x 64] . X^T      4 MMAs kind::f16, M=128 N=128 K=16, A = Theta in TMEM
//   R^T  = score(y, eta)                        epilogue warps; packed fp16 pairs back into TMEM
//   G^T[S x 64]   += R^T[S x 128] . X           8 MMAs kind::f16, M=64 N=64 K=16, A = R^T in TMEM
//
// What fp16 costs: range, not precision. (1) Theta is scaled by a power of two 2^k so that its
// largest magnitude lands in [2^10, 2^11) before it is split into hi = fp16(theta 2^k) and
// lo = fp16(theta 2^k - hi) (two MMA rows per particle exactly as in dense_tc.cuh: 22 significand
// bits, no systematic per-particle rounding error); the epilogue multiplies eta by 2^-k. (2) X must
// fit: the host selects this kernel only after checking the design matrix once per plan
// (engine/plan.py: largest magnitude below 2^15, no column whose root-mean-square is so small
// that fp16's absolute floor 2^-25 would matter), and the conversion warps check every element
// of every step: an entry outside the range raises MNF_ST_RANGE (no silent wrong answer; the
// caller switches to dense_precision="tf32"). (3) scores that overflow fp16 make the gradient
// non-finite, which MNF_ST_NONFINITE reports.
//
// Precision mode (stated): X and the scores rounded to nearest-even to fp16 (11-bit significand,
// unbiased, averages out over rows), theta as hi + lo fp16 pairs (22 bits), exact products, fp32
// accumulation in TMEM drained every kFlush tiles, fp32 / fp64 SIMT log-densities and sums.
//
// Warps (608 threads): 0-7 epilogue (two per TMEM lane quadrant, half a tile each), 8 TMA producer of
// the staging ring, 9 responses, 10 TMEM allocation + MMAs, 11-18 conversion (16 tile rows each, two
// threads per row).
//
// Replaces the same reference code as dense_tc.cuh: aten::mv / MvBackward of `X @ theta`
// (tests/test_mininf.py:11, examples/minibatch.md:33) and the Normal / Bernoulli / Poisson log_prob
// chains with their autograd twins (mininf/core.py:241), for all S particles in one pass over X.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "dense_simt.cuh"
#include "dense_tc.cuh"

namespace mnf {
namespace th {

using tc::elect_one;
using tc::fence_proxy_async;
using tc::lds128;
using tc::mbar_arrive;
using tc::mbar_arrive_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_desc;
using tc::smem_u32;
using tc::sts128;
using tc::tc_commit;
using tc::tc_fence_after;
using tc::tc_fence_before;
using tc::tc_ld32;
using tc::tc_ld_16x256b_x4;
using tc::tc_st16;
using tc::tc_wait_ld;
using tc::tc_wait_st;
using tc::tma_load_2d;
using tc::TileCounters;
#ifdef MNF_TC_DEBUG
using tc::g_tc_debug;
#endif

constexpr int kP = 64;
constexpr int kNS = 64;          // particle slots
constexpr int kMmaM = 2 * kNS;   // eta product: a hi and a lo row per particle
constexpr int kTileM = 128;
constexpr int kStgStages = 4;    // fp32 staging ring (32 KB per stage): three in flight while one is converted
constexpr int kImgStages = 4;    // fp16 operand images (16 KB per stage), live from eta(k) to G(k)
constexpr int kFlush = 8;
#ifndef MNF_TH_EPI_WARPS
#define MNF_TH_EPI_WARPS 8
#endif
// timing experiments only (wrong results): 1 no conversion, 2 no score math, 4 no gradient MMAs,
// 8 no eta MMAs, 16 no tcgen05.ld in the epilogue
#ifndef MNF_TH_DEV_SKIP
#define MNF_TH_DEV_SKIP 0
#endif
constexpr int kEpiWarps = MNF_TH_EPI_WARPS;   // warp w: TMEM lane quadrant w % 4, 32-row chunks (w / 4) kChunks .. + kChunks - 1
constexpr int kChunks = 16 / kEpiWarps;       // 32-row chunks of a tile scored by one epilogue warp
constexpr int kWarpTma = kEpiWarps, kWarpY = kEpiWarps + 1, kMmaWarp = kEpiWarps + 2;
constexpr int kFirstConv = kEpiWarps + 3;
constexpr int kConvWarps = 8;    // 16 tile rows each, two threads per row
constexpr int kThreads = (kFirstConv + kConvWarps) * 32;
static_assert(kEpiWarps == 4 || kEpiWarps == 8 || kEpiWarps == 16, "epilogue warps per TMEM lane quadrant: 1, 2 or 4");

constexpr uint32_t kAtomBytes = kTileM * 128;            // 128 rows x 32 fp32
constexpr uint32_t kStgBytes = 2 * kAtomBytes;           // fp32 tile: two 32-feature atoms
constexpr uint32_t kImgBytes = kTileM * 128;             // fp16 tile: 128 rows x 64 halves
constexpr uint32_t kYBytes = 2 * kTileM * 4;             // two planes of one float per row (response | liveness factor), permuted per 32-row chunk

constexpr uint32_t kOffStg = 0;
constexpr uint32_t kOffImg = kOffStg + kStgStages * kStgBytes;
constexpr uint32_t kOffY = kOffImg + kImgStages * kImgBytes;
constexpr uint32_t kOffBar = kOffY + kImgStages * kYBytes;
constexpr uint32_t kNumBars = 2 * kStgStages + 2 * kImgStages + 8;
constexpr uint32_t kOffMisc = kOffBar + 8 * kNumBars + (8 * kNumBars % 16 ? 8 : 0);
constexpr uint32_t kOffGrad = kOffMisc + 64 + kNS * 16;
constexpr uint32_t kOffStat = kOffGrad + kNS * (kP + 1) * 4;
constexpr uint32_t kSmemBytes = kOffStat + (kEpiWarps / 4) * 2 * kNS * 4 + 1024 /* alignment slack */;
static_assert(kOffMisc % 16 == 0, "misc block alignment");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

// tensor memory: two eta^T / R^T tiles, two gradient tiles, Theta (fp16 pairs: 32 columns)
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColEta = 0;                  // + b * kTileM
constexpr uint32_t kColG = 2 * kTileM;           // + gb * kP
constexpr uint32_t kColTheta = kColG + 2 * kP;   // 32 columns

// Instruction descriptor, kind::f16 with fp16 operands and fp32 accumulate:
//   [4,6) D format 1=F32 | [7,10) A format 0=F16 | [10,13) B format 0=F16 | [15] A MN-major
//   [16] B MN-major | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] . B[smem descriptor], kind::f16
__device__ __forceinline__ void mma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// two floats -> packed fp16 pair, `first` in the low half (the smaller k index)
__device__ __forceinline__ uint32_t pack_f16(float first, float second) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(second), "f"(first));
  return r;
}
__device__ __forceinline__ float f16_lo_as_float(uint32_t packed) {
  return __half2float(__ushort_as_half((unsigned short)(packed & 0xFFFFu)));
}
__device__ __forceinline__ float f16_hi_as_float(uint32_t packed) {
  return __half2float(__ushort_as_half((unsigned short)(packed >> 16)));
}
// 16 lanes x 16 columns as 8 registers per thread: register 2g + j of thread t holds lane
// t/4 + 8j, column 4g + t%4 (tools/tmem_shape_probe.cu)
__device__ __forceinline__ void tc_st_16x128b_x4(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.16x128b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// partial layout per CTA: [S][ncol], ncol = 1 + kP + 2 (same as the other dense kernels)
template <int FAMILY, bool ICPT>
__global__ void __launch_bounds__(kThreads, 1)
dense_th_kernel(const __grid_constant__ CUtensorMap map_x, mnf_dense_site_t site, const float* __restrict__ z,
                int S, int D, float* __restrict__ partial, uint32_t* __restrict__ status) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzled images need 1024-byte alignment
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sStg = base + kOffStg, sImg = base + kOffImg, sY = base + kOffY;
  const uint32_t bars = base + kOffBar;
  const uint32_t bStgFull = bars, bStgEmpty = bStgFull + 8 * kStgStages;
  const uint32_t bImgFull = bStgEmpty + 8 * kStgStages, bImgEmpty = bImgFull + 8 * kImgStages;
  const uint32_t bEtaFull = bImgEmpty + 8 * kImgStages, bRReady = bEtaFull + 16;
  const uint32_t bGFull = bRReady + 16, bGEmpty = bGFull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + kOffMisc);
  float* theta_scale = reinterpret_cast<float*>(gbase + kOffMisc + 8);   // [0] 2^k, [1] 2^-k
  TileCounters* counters = reinterpret_cast<TileCounters*>(gbase + kOffMisc + 16);
  DenseParticle* sPar = reinterpret_cast<DenseParticle*>(gbase + kOffMisc + 64);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (site.n_rows + kTileM - 1) / kTileM;
  const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  // ---- one-time setup ----------------------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < kStgStages; ++i) {
      mbar_init(bStgFull + 8 * i, 1);    // arrive.expect_tx of the TMA producer
      mbar_init(bStgEmpty + 8 * i, kConvWarps);   // one arrival per conversion warp
    }
    for (int i = 0; i < kImgStages; ++i) {
      mbar_init(bImgFull + 8 * i, kConvWarps + 1);    // the conversion warps + the y warp
      mbar_init(bImgEmpty + 8 * i, 1);   // tcgen05.commit after the gradient product
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bEtaFull + 8 * i, 1);
      mbar_init(bRReady + 8 * i, kEpiWarps * 32);
      mbar_init(bGFull + 8 * i, 1);
      mbar_init(bGEmpty + 8 * i, kEpiWarps * 32);
    }
    counters->n_live = 0.f;
    counters->lgamma_sum = 0.0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == kWarpTma && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
  for (int s = tid; s < kNS; s += kThreads) {
    DenseParticle dp;
    dp.icpt = 0.f; dp.scale = 1.f; dp.dscale = 0.f;
    if (s < S) {
      dp = dense_particle(site, z + (int64_t)s * D);
      if (FAMILY == MNF_NORMAL && !(dp.scale > 0.0f)) atomicOr(status, MNF_ST_BAD_PARAM);
    }
    sPar[s] = dp;
  }
  // power-of-two scale of Theta: the largest |theta| of any particle lands in [2^10, 2^11), so the
  // lo halves stay far above fp16's subnormal range (same arithmetic in every CTA)
  if (warp == 0) {
    float big = 0.f;
    for (int i = lane; i < S * kP; i += 32) big = fmaxf(big, fabsf(z[(int64_t)(i / kP) * D + site.theta_lat + i % kP]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) big = fmaxf(big, __shfl_xor_sync(0xffffffffu, big, o));
    if (lane == 0) {
      int e = 0;
      if (big > 0.f && big < 3.0e38f) frexpf(big, &e);      // big = m 2^e, m in [0.5, 1)
      const int k = max(-100, min(100, 11 - e));
      theta_scale[0] = ldexpf(1.0f, k);
      theta_scale[1] = ldexpf(1.0f, -k);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const float up = theta_scale[0], down = theta_scale[1];

  if (warp < 4) {
    // Theta -> TMEM as the A operand of the eta product (M = 128: MMA row m lives on lane m).
    // Lane 32*warp + i (i < 16) holds hi = fp16(theta 2^k) of particle 16*warp + i, lane
    // 32*warp + 16 + i its lo = fp16(theta 2^k - hi); features 2c, 2c + 1 packed in column
    // kColTheta + c. Spare particle slots are zero.
    const int s = warp * 16 + (lane & 15);
    const bool owner = s < S;
    const bool lo_row = lane >= 16;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll
    for (int ch = 0; ch < kP / 32; ++ch) {
      uint32_t v[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const int f = ch * 32 + 2 * c;
        const float t0 = owner ? z[(int64_t)s * D + site.theta_lat + f] * up : 0.f;
        const float t1 = owner ? z[(int64_t)s * D + site.theta_lat + f + 1] * up : 0.f;
        const uint32_t hi = pack_f16(t0, t1);
        v[c] = lo_row ? pack_f16(t0 - f16_lo_as_float(hi), t1 - f16_hi_as_float(hi)) : hi;
      }
      tc_st16(tmem + lane_base + kColTheta + ch * 16, v);
    }
    tc_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  if (warp == kWarpTma) {
    // ================= TMA producer of the fp32 staging ring (one elected lane) ================
    if (elect_one()) {
      for (int64_t k = 0; k < my_tiles; ++k) {
        const int st = (int)(k % kStgStages);
        const int row0 = (int)((blockIdx.x + k * gridDim.x) * kTileM);
        mbar_wait(bStgEmpty + 8 * st, (uint32_t)(((k / kStgStages) & 1) ^ 1));
        mbar_arrive_expect_tx(bStgFull + 8 * st, kStgBytes);
#pragma unroll
        for (int a = 0; a < kP / 32; ++a)
          tma_load_2d(sStg + (uint32_t)st * kStgBytes + a * kAtomBytes, &map_x, a * 32, row0, bStgFull + 8 * st);
      }
    }
    __syncwarp();
  } else if (warp >= kFirstConv) {
    // ================= conversion warps: fp32 staging -> fp16 operand image =====================
    // Two threads per row (a thread converts one 32-feature atom of its row: 8 LDS.128 -> 4 STS.128),
    // a warp covers 16 rows, eight warps a tile: the stage's latency (a chain of shared-memory round
    // trips) is on the critical path of every tile, so it is spread over many short threads.
    // Staging: per 32-feature atom, row r is 128 B with its 16-byte chunks XOR-ed with r % 8 (TMA
    // SWIZZLE_128B); image: row r is 128 B (64 halves) under the same swizzle, which is at once the
    // K-major operand of the eta product and the MN-major operand of the gradient product. The two
    // threads of a row walk their chunks in opposite halves (c ^ 4h), so the eight lanes of a
    // quarter-warp always touch eight different 16-byte bank groups.
    const uint32_t r = (uint32_t)(16 * (warp - kFirstConv) + (lane >> 1));
    const uint32_t h = (uint32_t)lane & 1u;
    const uint32_t sw = r & 7u;
    float worst = 0.f;                       // largest |x| seen (NaN entries surface as a NaN loss)
    TC_DECL();
    for (int64_t k = 0; k < my_tiles; ++k) {
      TC_T0();
      const int sst = (int)(k % kStgStages), ist = (int)(k % kImgStages);
      mbar_wait(bImgEmpty + 8 * ist, (uint32_t)(((k / kImgStages) & 1) ^ 1));
      mbar_wait(bStgFull + 8 * sst, (uint32_t)((k / kStgStages) & 1));
      TC_ACC(0);   // conv: wait
      const uint32_t src = sStg + (uint32_t)sst * kStgBytes + h * kAtomBytes + r * 128u;
      const uint32_t dst = sImg + (uint32_t)ist * kImgBytes + r * 128u;
      float4 x[8];
      if (!(MNF_TH_DEV_SKIP & 1)) {
#pragma unroll
      for (int c = 0; c < 8; ++c) x[c] = lds128(src + ((((uint32_t)c ^ (h << 2)) ^ sw) << 4));   // x[c] = chunk c ^ 4h
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // image chunk 4h + j holds features 32h + 8j .. + 7 = staging chunks 2j, 2j + 1 of atom h;
        // step i handles j = i ^ 2h (again opposite halves for the two threads of a row)
        const uint32_t j = (uint32_t)i ^ (h << 1);
        const float4 lo4 = x[(2 * i) ^ 0], hi4 = x[(2 * i) ^ 1];   // staging chunks (2i) ^ 4h, (2i + 1) ^ 4h = 2j, 2j + 1
        worst = fmaxf(worst, fmaxf(fmaxf(fabsf(lo4.x), fabsf(lo4.y)), fmaxf(fabsf(lo4.z), fabsf(lo4.w))));
        worst = fmaxf(worst, fmaxf(fmaxf(fabsf(hi4.x), fabsf(hi4.y)), fmaxf(fabsf(hi4.z), fabsf(hi4.w))));
        sts128(dst + (((4u * h + j) ^ sw) << 4), pack_f16(lo4.x, lo4.y), pack_f16(lo4.z, lo4.w), pack_f16(hi4.x, hi4.y),
               pack_f16(hi4.z, hi4.w));
      }
      }
      fence_proxy_async();                   // generic-proxy stores -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bImgFull + 8 * ist);
        mbar_arrive(bStgEmpty + 8 * sst);
      }
      TC_ACC(1);   // conv: convert
    }
    TC_FLUSH(10, 2, warp == kFirstConv && lane == 0);
    // anything at or above 2^15 (or infinite) is outside the range this operand format is used for
    if (!(worst < 32768.0f)) atomicOr(status, MNF_ST_RANGE);
  } else if (warp == kWarpY) {
    // ================= y warp: (y, live) pairs of each tile, four rows per lane ===============
    const bool y_vec = (reinterpret_cast<uintptr_t>(site.y) % 16 == 0) &&
                       (site.mask == nullptr || reinterpret_cast<uintptr_t>(site.mask) % 4 == 0);
    float4 yq[kImgStages];
    uint32_t mq[kImgStages];
    auto fetch = [&](int64_t k, float4& yraw, uint32_t& mraw) {
      const int64_t row = (blockIdx.x + k * gridDim.x) * kTileM + lane * 4;
      if (y_vec && row + 4 <= site.n_rows) {
        yraw = __ldg(reinterpret_cast<const float4*>(site.y + row));
        mraw = site.mask == nullptr ? 0x01010101u : __ldg(reinterpret_cast<const uint32_t*>(site.mask + row));
      } else {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        mraw = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (row + q < site.n_rows) {
            t[q] = __ldg(site.y + row + q);
            const uint32_t m = site.mask == nullptr ? 1u : (uint32_t)__ldg(site.mask + row + q);
            mraw |= (m != 0 ? 1u : 0u) << (8 * q);
          }
        }
        yraw = make_float4(t[0], t[1], t[2], t[3]);
      }
    };
#pragma unroll
    for (int i = 0; i < kImgStages; ++i) {
      yq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      mq[i] = 0;
      if (i < my_tiles) fetch(i, yq[i], mq[i]);
    }
    bool bad_value = false;
    int live_total = 0;
    double lgam_total = 0.0;
    for (int64_t k0 = 0; k0 < my_tiles; k0 += kImgStages) {
#pragma unroll
      for (int i = 0; i < kImgStages; ++i) {
        const int64_t k = k0 + i;
        if (k < my_tiles) {
          mbar_wait(bImgEmpty + 8 * i, (uint32_t)(((k / kImgStages) & 1) ^ 1));
          const float yr[4] = {yq[i].x, yq[i].y, yq[i].z, yq[i].w};
          float yv[4], av[4];
          float lgam = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool live = ((mq[i] >> (8 * q)) & 0xFFu) != 0;
            // plane 0: the response (0 on masked-out rows); plane 1: what the epilogue multiplies eta by -
            // Normal: -live 2^-k (the power-of-two scale of Theta folded in, so that a score is two
            // FFMAs: y - 2^-k (hi + lo)); other families: live
            yv[q] = live ? yr[q] : 0.f;
            av[q] = live ? (FAMILY == MNF_NORMAL ? -down : 1.0f) : 0.f;
            if (live) {
              ++live_total;
              if (!in_support(FAMILY, yr[q])) bad_value = true;     // a live NaN is reported, not scored
              if (FAMILY == MNF_POISSON) lgam += log_factorial(yr[q]);
            }
          }
          if (FAMILY == MNF_POISSON) lgam_total += (double)lgam;
          // Shared-memory bandwidth is this kernel's scarcest resource, so the responses are laid out
          // for the epilogue's access pattern: within a 32-row chunk, row 8g + 2q + j is stored at
          // position 8q + 2g + j - the eight rows an epilogue thread (t % 4 == q) scores are then two
          // 16-byte words. This lane holds rows 4*lane .. 4*lane + 3 of the tile.
          const int rc = 4 * (lane & 7);                          // first row within the chunk
          const int g0 = rc >> 3, q0 = (rc & 7) >> 1;             // rows rc, rc+1 -> (g0, q0); rc+2, rc+3 -> (g0, q0 + 1)
          const uint32_t chunk_base = sY + (uint32_t)i * kYBytes + (uint32_t)(lane >> 3) * 128u;
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(chunk_base + (uint32_t)(8 * q0 + 2 * g0) * 4u), "f"(yv[0]), "f"(yv[1]) : "memory");
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(chunk_base + (uint32_t)(8 * (q0 + 1) + 2 * g0) * 4u), "f"(yv[2]), "f"(yv[3]) : "memory");
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(chunk_base + kTileM * 4u + (uint32_t)(8 * q0 + 2 * g0) * 4u), "f"(av[0]), "f"(av[1]) : "memory");
          asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(chunk_base + kTileM * 4u + (uint32_t)(8 * (q0 + 1) + 2 * g0) * 4u), "f"(av[2]), "f"(av[3]) : "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(bImgFull + 8 * i);
          if (k + kImgStages < my_tiles) fetch(k + kImgStages, yq[i], mq[i]);
        }
      }
    }
    if (bad_value) atomicOr(status, MNF_ST_BAD_VALUE);
    const float live_sum = warp_sum((float)live_total);
    if (FAMILY == MNF_POISSON) lgam_total = warp_sum(lgam_total);
    if (lane == 0) {
      counters->n_live = live_sum;
      counters->lgamma_sum = lgam_total;
    }
    asm volatile("bar.sync 1, %0;" ::"n"((kEpiWarps + 1) * 32) : "memory");
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer: warp-uniform loop, one elected lane issues ================
    constexpr uint32_t idesc_eta = idesc_f16(kMmaM, kTileM, 0, 0);   // M=128 N=128, B K-major
    constexpr uint32_t idesc_g = idesc_f16(kNS, kP, 0, 1);           // M=64  N=64,  B MN-major
    // one image, two views: rows of 128 B, SWIZZLE_128B, 8-row groups 1024 B apart
    const uint64_t dImg = smem_desc(sImg, 16, 1024, 2);
    const uint32_t d_lo = (uint32_t)dImg, d_hi = (uint32_t)(dImg >> 32);
    TC_DECL();
    for (int64_t k = 0; k <= my_tiles; ++k) {
      TC_T0();
      if (k < my_tiles) {
        const int ist = (int)(k % kImgStages);
        const uint32_t b = (uint32_t)(k & 1);
        mbar_wait(bImgFull + 8 * ist, (uint32_t)((k / kImgStages) & 1));
        TC_ACC(0);   // mma: wait operands
        tc_fence_after();
        if (elect_one()) {
          const uint32_t lo = d_lo + (uint32_t)ist * (kImgBytes >> 4);
          const uint32_t d = tmem + kColEta + b * kTileM;
#pragma unroll
          for (int ks = 0; ks < kP / 16; ++ks)      // 16 features = 32 B along the swizzled row
            if (!(MNF_TH_DEV_SKIP & 8)) mma_ts_f16(d, tmem + kColTheta + ks * 8, lo + ks * 2, d_hi, idesc_eta, ks != 0 ? 1u : 0u);
          tc_commit(bEtaFull + 8 * b);
        }
        __syncwarp();
        TC_ACC(1);   // mma: issue eta
      }
      if (k >= 1) {
        const int64_t kk = k - 1;
        const int ist = (int)(kk % kImgStages);
        const uint32_t b = (uint32_t)(kk & 1);
        const int64_t grp = kk / kFlush;
        const uint32_t gb = (uint32_t)(grp & 1);
        const bool first = (kk % kFlush) == 0;
        const bool last = (kk % kFlush) == kFlush - 1 || kk == my_tiles - 1;
        mbar_wait(bRReady + 8 * b, (uint32_t)((kk >> 1) & 1));
        TC_ACC(2);   // mma: wait r_ready
        if (first) mbar_wait(bGEmpty + 8 * gb, (uint32_t)(((grp >> 1) & 1) ^ 1));
        tc_fence_after();
        if (elect_one()) {
          const uint32_t lo = d_lo + (uint32_t)ist * (kImgBytes >> 4);
          const uint32_t d = tmem + kColG + gb * kP;
          const uint32_t a0 = tmem + kColEta + b * kTileM;
#pragma unroll
          for (int ks = 0; ks < kTileM / 16; ++ks) {
            // 16 tile rows per MMA: their packed scores are 8 TMEM columns (rows 32Q .. 32Q + 31 at
            // columns 32Q .. 32Q + 15 of the eta tile: every epilogue warp writes inside the columns
            // it read), their X rows 2048 B of the image
            const uint32_t a_col = (uint32_t)(32 * (ks >> 1) + 8 * (ks & 1));   // chunk ks / 2
            if (!(MNF_TH_DEV_SKIP & 4)) mma_ts_f16(d, a0 + a_col, lo + ks * (2048 >> 4), d_hi, idesc_g, (!first || ks > 0) ? 1u : 0u);
          }
          tc_commit(bImgEmpty + 8 * ist);
          if (last) tc_commit(bGFull + 8 * gb);
        }
        __syncwarp();
        TC_ACC(3);   // mma: issue G
      }
    }
    TC_FLUSH(4, 4, lane == 0);
  } else if (warp < kEpiWarps) {
    // ================= epilogue warps: quadrant q = warp % 4, chunks Q kChunks .. of every tile, Q = warp / 4 ====
    // The score stage is bound by instruction issue (8192 points per tile over 128 lanes: every
    // instruction per point is 64 issue cycles per tile and scheduler, of ~1000-1100 available at the
    // power-capped clock), so the per-point work of the Normal family is three FFMAs and half a
    // conversion: score = y + nl (hi + lo) as two FFMAs with nl = -live 2^-k staged per row by the
    // response warp, sum r^2 as one FFMA, two scores per cvt.rn.f16x2.
    const int q = warp & 3, Q = warp >> 2;
    const int s = q * 16 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    double stA_total = 0.0, stB_total = 0.0;  // Normal: sum r^2 | others: sum log-density (w/o lgamma)
    double scA_total = 0.0, scB_total = 0.0;  // sum of scores: the intercept gradient (ICPT only)
    const float icptA = ICPT ? sPar[q * 16 + (lane >> 2)].icpt : 0.f;
    const float icptB = ICPT ? sPar[q * 16 + (lane >> 2) + 8].icpt : 0.f;
    constexpr int kDrainCols = kP / (kEpiWarps / 4);   // features drained by this warp
    float* grad_row = reinterpret_cast<float*>(gbase + kOffGrad) + (size_t)(lane < 16 ? s : 0) * (kP + 1) + kDrainCols * Q;
    if (lane < 16)
      for (int j = 0; j < kDrainCols; ++j) grad_row[j] = 0.f;
    int64_t n_drained = 0;

    auto drain = [&]() {
      const int64_t grp = n_drained;
      const uint32_t gb = (uint32_t)(grp & 1);
      mbar_wait(bGFull + 8 * gb, (uint32_t)((grp >> 1) & 1));
      tc_fence_after();
      // 32x32b: thread == TMEM lane; an M = 64 accumulator lives on lanes 0-15
#pragma unroll
      for (int c0 = 0; c0 < kDrainCols; c0 += 16) {
        uint32_t v[16];
        tc::tc_ld16(tmem + lane_base + kColG + gb * kP + Q * kDrainCols + c0, v);
        tc_wait_ld();
        if (lane < 16) {
#pragma unroll
          for (int c = 0; c < 16; ++c) grad_row[c0 + c] += __uint_as_float(v[c]);
        }
      }
      tc_fence_before();
      mbar_arrive(bGEmpty + 8 * gb);
      ++n_drained;
    };

    // one (row, particle) point; `aux` is the row's staged factor (Normal: -live 2^-k, others: live)
    const float icptUpA = icptA * up, icptUpB = icptB * up;   // intercept in units of the scaled eta (2^k is exact)
    auto point = [&](uint32_t cell, uint32_t cell_lo, float y, float aux, float icpt, float icpt_up, float& stat, float& ssum) {
      float score;
      if (FAMILY == MNF_NORMAL) {
        // live * (y - icpt - 2^-k (hi + lo)); 1/sigma^2 applied at the end
        const float t = ICPT ? fmaf(aux, icpt_up, y) : y;
        score = fmaf(aux, __uint_as_float(cell), fmaf(aux, __uint_as_float(cell_lo), t));
        stat = fmaf(score, score, stat);
      } else {
        const float eta = ICPT ? fmaf(__uint_as_float(cell) + __uint_as_float(cell_lo), down, icpt)
                               : (__uint_as_float(cell) + __uint_as_float(cell_lo)) * down;
        if (FAMILY == MNF_BERNOULLI_LOGITS) {
          const float e = __expf(-fabsf(eta));
          const float inv = __fdividef(1.0f, 1.0f + e);
          const float sig = eta >= 0.f ? inv : e * inv;
          score = aux * (y - sig);
          stat += aux * (y * eta - (fmaxf(eta, 0.f) + __logf(1.0f + e)));   // softplus, abs. error ~1e-7
        } else {
          const float rate = __expf(eta);
          score = aux * (y - rate);
          stat += aux * fmaf(y, eta, -rate);
        }
      }
      if (ICPT) ssum += score;
      return score;
    };
    // 32 columns (tile rows) -> 16 packed columns: thread t touches rows 32ch + 8g + 2(t%4) + {0,1}
    // of particles A = 16q + t/4 and B = A + 8; a pair of rows is one fp16x2 word of the A operand
    auto process = [&](const uint32_t (&v)[16], const uint32_t (&l)[16], uint32_t (&w)[8], const float4* yl, int ch,
                       float& sa, float& sb, float& ra, float& rb) {
      // the eight responses and factors this thread needs: two 16-byte words each (see the y warp's layout)
      const float4 y0 = yl[8 * ch + 2 * (lane & 3)], y1 = yl[8 * ch + 2 * (lane & 3) + 1];
      const float4 a0 = yl[kTileM / 4 + 8 * ch + 2 * (lane & 3)], a1 = yl[kTileM / 4 + 8 * ch + 2 * (lane & 3) + 1];
      const float yv[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float pa0 = point(v[4 * g + 0], l[4 * g + 0], yv[2 * g], av[2 * g], icptA, icptUpA, sa, ra);
        const float pa1 = point(v[4 * g + 1], l[4 * g + 1], yv[2 * g + 1], av[2 * g + 1], icptA, icptUpA, sa, ra);
        const float pb0 = point(v[4 * g + 2], l[4 * g + 2], yv[2 * g], av[2 * g], icptB, icptUpB, sb, rb);
        const float pb1 = point(v[4 * g + 3], l[4 * g + 3], yv[2 * g + 1], av[2 * g + 1], icptB, icptUpB, sb, rb);
        w[2 * g + 0] = pack_f16(pa0, pa1);
        w[2 * g + 1] = pack_f16(pb0, pb1);
      }
    };

    constexpr uint32_t kLoRows = 16u << 16;   // TMEM lane offset of the lo rows inside a quadrant
    TC_DECL();
    for (int64_t k = 0; k < my_tiles; ++k) {
      TC_T0();
      const uint32_t b = (uint32_t)(k & 1);
      const float4* yl = reinterpret_cast<const float4*>(gbase + kOffY + (size_t)(k % kImgStages) * kYBytes);
      // the gradient group that ended two tiles ago has been issued (its R tile was handed over
      // two iterations back), so waiting for its commit cannot deadlock
      if (k >= 2 && ((k - 2) % kFlush) == kFlush - 1) drain();
      mbar_wait(bEtaFull + 8 * b, (uint32_t)((k >> 1) & 1));
      TC_ACC(0);   // epi: wait eta_full (+ drain)
      tc_fence_after();
      // this warp's chunks of the eta tile: 32 columns (tile rows) each, hi rows on lanes 0-15 and lo
      // rows on lanes 16-31 of the quadrant; the packed scores of a chunk go back into its first 16
      // columns (loaded before they are overwritten; no other warp touches a chunk's columns)
      const uint32_t t_eta = tmem + lane_base + kColEta + b * kTileM + 32 * kChunks * Q;
      float sa = 0.f, sb = 0.f, ra = 0.f, rb = 0.f;
      uint32_t v[kChunks][16], l[kChunks][16];
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        if (MNF_TH_DEV_SKIP & 16) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[c][i] = l[c][i] = (uint32_t)k;
        } else {
          tc_ld_16x256b_x4(t_eta + 32 * c, v[c]);
          tc_ld_16x256b_x4(t_eta + kLoRows + 32 * c, l[c]);
        }
      }
      tc_wait_ld();
      TC_ACC(1);   // epi: tcgen05.ld
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        uint32_t w[8];
        if (MNF_TH_DEV_SKIP & 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) w[i] = v[c][i] ^ l[c][i];
        } else {
          process(v[c], l[c], w, yl, kChunks * Q + c, sa, sb, ra, rb);
        }
        tc_st_16x128b_x4(t_eta + 32 * c, w);
      }
      TC_ACC(2);   // epi: scores
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(bRReady + 8 * b);
      stA_total += (double)sa;
      stB_total += (double)sb;
      if (ICPT) {
        scA_total += (double)ra;
        scB_total += (double)rb;
      }
      TC_ACC(3);   // epi: wait st, arrive
    }
    TC_FLUSH(0, 4, tid == 0);
    {
      const int64_t n_grp = (my_tiles + kFlush - 1) / kFlush;
      while (n_drained < n_grp) drain();
    }
    // the four threads t%4 = 0..3 of a quad hold partial sums of the same two particles
    stA_total += __shfl_xor_sync(0xffffffffu, stA_total, 1);
    stA_total += __shfl_xor_sync(0xffffffffu, stA_total, 2);
    stB_total += __shfl_xor_sync(0xffffffffu, stB_total, 1);
    stB_total += __shfl_xor_sync(0xffffffffu, stB_total, 2);
    if (ICPT) {
      scA_total += __shfl_xor_sync(0xffffffffu, scA_total, 1);
      scA_total += __shfl_xor_sync(0xffffffffu, scA_total, 2);
      scB_total += __shfl_xor_sync(0xffffffffu, scB_total, 1);
      scB_total += __shfl_xor_sync(0xffffffffu, scB_total, 2);
    }
    float* s_stat = reinterpret_cast<float*>(gbase + kOffStat) + Q * 2 * kNS;   // [kEpiWarps / 4][2][kNS]
    if ((lane & 3) == 0) {
      s_stat[q * 16 + (lane >> 2)] = (float)stA_total;
      s_stat[q * 16 + (lane >> 2) + 8] = (float)stB_total;
      if (ICPT) {
        s_stat[kNS + q * 16 + (lane >> 2)] = (float)scA_total;
        s_stat[kNS + q * 16 + (lane >> 2) + 8] = (float)scB_total;
      }
    }
    // every epilogue warp's statistics and gradient columns, and the y warp's counters, are final
    asm volatile("bar.sync 1, %0;" ::"n"((kEpiWarps + 1) * 32) : "memory");
    const float* s_all = reinterpret_cast<const float*>(gbase + kOffStat);
    const bool owner_thread = Q == 0 && lane < 16;
    float st0 = 0.f, sc0 = 0.f;
    if (owner_thread) {
#pragma unroll
      for (int i = 0; i < kEpiWarps / 4; ++i) {
        st0 += s_all[2 * i * kNS + s];
        if (ICPT) sc0 += s_all[(2 * i + 1) * kNS + s];
      }
    }
    grad_row -= kDrainCols * Q;

    // ---- per-particle results: this thread is the only owner of particle s --------------------
    if (owner_thread && s < S) {
      const int ncol = 1 + kP + 2;
      float* out = partial + ((size_t)blockIdx.x * S + s) * ncol;
      const DenseParticle pp = sPar[s];
      const float cnt = *reinterpret_cast<volatile float*>(&counters->n_live);
      const double lgsum = *reinterpret_cast<volatile double*>(&counters->lgamma_sum);
      float lp, gscale = 1.0f, dscale = 0.f;
      if (FAMILY == MNF_NORMAL) {
        const float inv = 1.0f / pp.scale, iv = inv * inv;
        lp = -0.5f * iv * st0 - cnt * (logf(pp.scale) + kLogSqrt2Pi);
        dscale = (st0 * iv * inv - cnt * inv) * pp.dscale;
        gscale = iv;
      } else if (FAMILY == MNF_BERNOULLI_LOGITS) {
        lp = st0;
      } else {
        lp = st0 - (float)lgsum;
      }
      out[0] = lp;
#pragma unroll
      for (int j = 0; j < kP; ++j) out[1 + j] = grad_row[j] * gscale;
      out[1 + kP] = sc0 * gscale;   // intercept gradient
      out[2 + kP] = dscale;
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
  }
}

}  // namespace th
}  // namespace mnf
