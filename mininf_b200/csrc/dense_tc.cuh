// Dense-link sweep, tcgen05 variant (MNF_DENSE_TF32) for sm_100a.
//
// One persistent CTA per SM streams 128-row tiles of X exactly once and runs BOTH matrix products
// of the step on the 5th-generation tensor cores, with the X tile shared between them in smem:
//
//   eta[128 x S]  = Xtile[128 x p] . Theta^T[p x S]       tcgen05.mma kind::tf32, M=128 N=S K=p
//                   A = Xtile as a K-major  SWIZZLE_128B operand, B = Theta K-major, D in TMEM
//   R  [128 x S]  = score(y, eta)                          epilogue warps: tcgen05.ld -> registers
//   G  [p x S]   += Xtile^T[p x 128] . R[128 x S]          tcgen05.mma kind::tf32, M=p N=S K=128
//                   A = the SAME smem tile read as an MN-major operand, B = R MN-major,
//                   D accumulates in TMEM across every tile the CTA owns
//
// Precision mode (stated): operands are rounded to nearest to TF32 (10-bit mantissa)
// IN THE KERNEL — X by the producer warps on its way from HBM to shared memory, Theta and R before
// they are stored — products are exact, accumulation is fp32 in TMEM. The tensor core itself
// truncates fp32 inputs, which would bias eta by ~2^-11 relative; rounding first removes the bias
// (measured effect in DESIGN.md). Log-densities, residual statistics and all reductions are fp32
// / fp64 SIMT.
//
// Warp roles (416 threads): warps 0-3 epilogue (TMEM lane quadrant == warp id), warps 4-11 two
// producer groups that alternate tiles (coalesced 128-bit LDG -> round -> swizzled STS.128 ->
// fence.proxy.async -> mbarrier), warp 12 allocates TMEM and one elected lane issues every MMA.
// Pipelines: X ring (kStages), double-buffered eta accumulator and R tile, all mbarrier based.
//
// Replaces: aten::mv / addmv_ and MvBackward of `X @ theta` (tests/test_mininf.py:11,
// examples/minibatch.md:33) plus the element-wise Normal / Bernoulli / Poisson log_prob chains
// and their autograd twins (mininf/core.py:241), for all S particles in one pass.
#pragma once

#include "common.cuh"
#include "dense_simt.cuh"

namespace mnf {
namespace tc {

constexpr int kP = 64;          // features handled by this instantiation
constexpr int kNS = 64;         // particle slots (MMA N); S <= kNS, spare slots carry theta = 0
constexpr int kTileM = 128;     // rows per tile
constexpr int kStages = 4;      // X ring depth
constexpr int kEpiWarps = 4;
constexpr int kProdWarps = 8;   // two groups of four
constexpr int kThreads = (kEpiWarps + kProdWarps + 1) * 32;
constexpr int kMmaWarp = kEpiWarps + kProdWarps;

constexpr uint32_t kAtomBytes = kTileM * 128;                 // one K-atom of the X tile: 128 rows x 128 B
constexpr uint32_t kXStageBytes = (kP / 32) * kAtomBytes;     // 32 KB
constexpr uint32_t kRStageBytes = (kNS / 32) * kAtomBytes;    // 32 KB
constexpr uint32_t kThetaAtomBytes = kNS * 128;               // 8 KB
constexpr uint32_t kThetaBytes = (kP / 32) * kThetaAtomBytes; // 16 KB

constexpr uint32_t kOffX = 0;
constexpr uint32_t kOffR = kOffX + kStages * kXStageBytes;
constexpr uint32_t kOffTheta = kOffR + 2 * kRStageBytes;
constexpr uint32_t kOffBar = kOffTheta + kThetaBytes;
constexpr uint32_t kNumBars = 2 * kStages + 4 * 2 + 1;
constexpr uint32_t kOffMisc = kOffBar + 8 * kNumBars;         // tmem address + per-particle params
constexpr uint32_t kSmemBytes = kOffMisc + 16 + kNS * 16 + 1024 /* alignment slack */;

constexpr uint32_t kTmemCols = 256;   // eta0 [0,64) eta1 [64,128) G [128,192)
constexpr uint32_t kColEta = 0;
constexpr uint32_t kColG = 2 * kNS;

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns of TMEM -> 16 registers per thread
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors (cute/arch/mma_sm100_desc.hpp, cute/atom/mma_traits_sm100.hpp) -------------
// Shared-memory matrix descriptor, SWIZZLE_128B, version 1 (Blackwell):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 | [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// K-major operand: rows of 128 B (32 fp32 of K), 8-row groups 1024 B apart (SBO); LBO unused.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t addr) { return smem_desc(addr, 16, 1024); }
// MN-major operand: 32 fp32 of M/N contiguous (128 B), K rows 128 B apart, 8-row K groups 1024 B
// apart (SBO), next 32-element M/N chunk one atom (128 rows x 128 B) further (LBO).
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t addr) {
  return smem_desc(addr, kAtomBytes, 1024);
}
// Instruction descriptor, kind::tf32, fp32 accumulate:
//   [4,6) D format 1=F32 | [7,10) A format 2=TF32 | [10,13) B format 2=TF32 | [15] A MN-major
//   [16] B MN-major | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Swizzle<3,4,3>: XOR the 16-byte chunk index (bits 4-6) with the 128-byte row index (bits 7-9).
__device__ __forceinline__ uint32_t swz128(uint32_t byte_off) {
  return byte_off ^ (((byte_off >> 7) & 7u) << 4);
}
// round-to-nearest (ties away) to TF32: the tensor core then only drops zero bits
__device__ __forceinline__ uint32_t rn_tf32(float x) {
  return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
}
__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

struct ParticleParam {
  float icpt;
  float scale;
  float dscale;
  float pad;
};

// partial layout per CTA: [S][ncol], ncol = 1 + kP + 2 (same as the SIMT variant)
template <int FAMILY>
__global__ void __launch_bounds__(kThreads, 1)
dense_tc_kernel(mnf_dense_site_t site, const float* __restrict__ z, int S, int D,
                float* __restrict__ partial, uint32_t* __restrict__ status) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operands need 1024-byte alignment
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sX = base + kOffX, sR = base + kOffR, sTheta = base + kOffTheta;
  const uint32_t bars = base + kOffBar;
  // barrier map
  const uint32_t bXFull = bars, bXEmpty = bars + 8 * kStages;
  const uint32_t bEtaFull = bars + 16 * kStages, bEtaEmpty = bEtaFull + 16;
  const uint32_t bRFull = bEtaEmpty + 16, bREmpty = bRFull + 16;
  const uint32_t bGFull = bREmpty + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + kOffMisc);
  ParticleParam* sPar = reinterpret_cast<ParticleParam*>(gbase + kOffMisc + 16);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  const int64_t n_tiles = (site.n_rows + kTileM - 1) / kTileM;
  // tiles owned by this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  // ---- one-time setup ----------------------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bXFull + 8 * i, 4);   // four producer warps of the owning group
      mbar_init(bXEmpty + 8 * i, 1);  // tcgen05.commit
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bEtaFull + 8 * i, 1);
      mbar_init(bEtaEmpty + 8 * i, kEpiWarps * 32);
      mbar_init(bRFull + 8 * i, kEpiWarps * 32);
      mbar_init(bREmpty + 8 * i, 1);
    }
    mbar_init(bGFull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // Theta -> smem, K-major SWIZZLE_128B, rounded to TF32; spare particle slots are zero
  for (int i = tid; i < kNS * kP; i += kThreads) {
    const int s = i / kP, j = i % kP;
    const float v = s < S ? z[(int64_t)s * D + site.theta_lat + j] : 0.0f;
    const uint32_t off = (uint32_t)(j >> 5) * kThetaAtomBytes + swz128((uint32_t)s * 128u + (uint32_t)(j & 31) * 4u);
    *reinterpret_cast<uint32_t*>(gbase + kOffTheta + off) = rn_tf32(v);
  }
  for (int s = tid; s < kNS; s += kThreads) {
    ParticleParam pp;
    pp.icpt = 0.f; pp.scale = 1.f; pp.dscale = 0.f; pp.pad = 0.f;
    if (s < S) {
      const DenseParticle dp = dense_particle(site, z + (int64_t)s * D);
      pp.icpt = dp.icpt; pp.scale = dp.scale; pp.dscale = dp.dscale;
      if (FAMILY == MNF_NORMAL && !(dp.scale > 0.0f)) atomicOr(status, MNF_ST_BAD_PARAM);
    }
    sPar[s] = pp;
  }
  fence_proxy_async();  // Theta was written through the generic proxy, tcgen05.mma reads it async
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= kEpiWarps && warp < kMmaWarp) {
    // ================= producers: HBM -> registers -> TF32 round -> swizzled smem ============
    const int group = (warp - kEpiWarps) >> 2;           // 0 or 1
    const int gtid = ((warp - kEpiWarps) & 3) * 32 + lane;  // 0..127 within the group
    const bool vec_ok = true;
    (void)vec_ok;
    for (int64_t k = group; k < my_tiles; k += 2) {
      const int64_t tile = blockIdx.x + k * gridDim.x;
      const int64_t row0 = tile * kTileM;
      const int st = (int)(k % kStages);
      const uint32_t use = (uint32_t)(k / kStages);
      float4 v[16];
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int f = it * 128 + gtid;   // float4 index in the tile: 16 per row
        const int r = f >> 4, c = f & 15;
        const int64_t row = row0 + r;
        if (row < site.n_rows) v[it] = ldg_stream(reinterpret_cast<const float4*>(site.X + row * site.ldx) + c);
        else v[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      mbar_wait(bXEmpty + 8 * st, (use & 1u) ^ 1u);
      const uint32_t stage = sX + (uint32_t)st * kXStageBytes;
#pragma unroll
      for (int it = 0; it < 16; ++it) {
        const int f = it * 128 + gtid;
        const uint32_t r = (uint32_t)(f >> 4), c = (uint32_t)(f & 15);
        const uint32_t off = (c >> 3) * kAtomBytes + r * 128u + (((c & 7u) ^ (r & 7u)) << 4);
        sts128(stage + off, rn_tf32(v[it].x), rn_tf32(v[it].y), rn_tf32(v[it].z), rn_tf32(v[it].w));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bXFull + 8 * st);
    }
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer (one elected lane) ==========================================
    if (lane == 0) {
      constexpr uint32_t idesc_eta = idesc_tf32(kTileM, kNS, 0, 0);
      constexpr uint32_t idesc_g = idesc_tf32(kP, kNS, 1, 1);
      for (int64_t k = 0; k <= my_tiles; ++k) {
        if (k < my_tiles) {
          const int st = (int)(k % kStages);
          const uint32_t b = (uint32_t)(k & 1);
          mbar_wait(bXFull + 8 * st, (uint32_t)((k / kStages) & 1));
          mbar_wait(bEtaEmpty + 8 * b, (uint32_t)(((k >> 1) & 1) ^ 1));
          tc_fence_after();
          const uint32_t xs = sX + (uint32_t)st * kXStageBytes;
#pragma unroll
          for (int a = 0; a < kP / 32; ++a) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              tc_mma_tf32(tmem + kColEta + b * kNS, desc_kmajor(xs + a * kAtomBytes + ks * 32),
                          desc_kmajor(sTheta + a * kThetaAtomBytes + ks * 32), idesc_eta,
                          (a | ks) != 0 ? 1u : 0u);
            }
          }
          tc_commit(bEtaFull + 8 * b);
        }
        if (k >= 1) {
          const int64_t kk = k - 1;
          const int st = (int)(kk % kStages);
          const uint32_t b = (uint32_t)(kk & 1);
          mbar_wait(bRFull + 8 * b, (uint32_t)((kk >> 1) & 1));
          tc_fence_after();
          const uint32_t xs = sX + (uint32_t)st * kXStageBytes;
          const uint32_t rs = sR + b * kRStageBytes;
#pragma unroll
          for (int ks = 0; ks < kTileM / 8; ++ks) {
            tc_mma_tf32(tmem + kColG, desc_mnmajor(xs + ks * 1024), desc_mnmajor(rs + ks * 1024),
                        idesc_g, (kk > 0 || ks > 0) ? 1u : 0u);
          }
          tc_commit(bXEmpty + 8 * st);
          tc_commit(bREmpty + 8 * b);
        }
      }
      tc_commit(bGFull);
    }
    __syncwarp();
  } else {
    // ================= epilogue warps: eta -> log-density, score R ============================
    const int row_in_tile = tid;  // TMEM lane == tile row
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float st0[kNS];  // Normal: sum r^2            | others: sum log-density
#pragma unroll
    for (int s = 0; s < kNS; ++s) st0[s] = 0.f;
    float n_live = 0.f;
    double lgam = 0.0;   // Poisson: sum lgamma(y+1) over live rows (particle independent)
    bool bad_value = false;

    // prefetch y / mask of the first tile
    float y_next = 0.f; bool live_next = false;
    {
      const int64_t row = (int64_t)blockIdx.x * kTileM + row_in_tile;
      live_next = row < site.n_rows && (site.mask == nullptr || site.mask[row] != 0);
      if (live_next) y_next = __ldg(site.y + row);
    }
    for (int64_t k = 0; k < my_tiles; ++k) {
      const uint32_t b = (uint32_t)(k & 1);
      const float y = y_next;
      const bool live = live_next;
      if (k + 1 < my_tiles) {
        const int64_t row = (blockIdx.x + (k + 1) * gridDim.x) * kTileM + row_in_tile;
        live_next = row < site.n_rows && (site.mask == nullptr || site.mask[row] != 0);
        y_next = live_next ? __ldg(site.y + row) : 0.f;
      }
      if (live) {
        n_live += 1.f;
        if (y != y) bad_value = true;
        if (FAMILY == MNF_POISSON) lgam += (double)lgammaf(y + 1.0f);
      }
      mbar_wait(bEtaFull + 8 * b, (uint32_t)((k >> 1) & 1));
      tc_fence_after();
      mbar_wait(bREmpty + 8 * b, (uint32_t)(((k >> 1) & 1) ^ 1));
      const uint32_t rs = sR + b * kRStageBytes;
      const uint32_t r = (uint32_t)row_in_tile;
#pragma unroll
      for (int ch = 0; ch < kNS / 16; ++ch) {
        uint32_t v[16];
        tc_ld16(tmem + lane_base + kColEta + b * kNS + ch * 16, v);
        tc_wait_ld();
        uint32_t out[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          const int s = ch * 16 + c;
          const float eta = __uint_as_float(v[c]);
          float score;
          if (FAMILY == MNF_NORMAL) {
            score = live ? y - eta : 0.f;          // unit-scale residual; 1/sigma^2 applied at the end
            st0[s] = fmaf(score, score, st0[s]);
          } else if (FAMILY == MNF_BERNOULLI_LOGITS) {
            const float e = __expf(-fabsf(eta));
            const float inv = __fdividef(1.0f, 1.0f + e);
            const float sig = eta >= 0.f ? inv : e * inv;
            const float lp = y * eta - (fmaxf(eta, 0.f) + log1pf(e));
            score = live ? y - sig : 0.f;
            st0[s] += live ? lp : 0.f;
          } else {
            const float rate = expf(eta);
            score = live ? y - rate : 0.f;
            st0[s] += live ? fmaf(y, eta, -rate) : 0.f;
          }
          out[c] = rn_tf32(score);
        }
        // R tile, MN-major SWIZZLE_128B: row r holds 32 particles per 128-byte line
        const uint32_t half = (uint32_t)(ch >> 1) * kAtomBytes + r * 128u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t c16 = (uint32_t)((ch & 1) * 4 + q);
          sts128(rs + half + ((c16 ^ (r & 7u)) << 4), out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(bEtaEmpty + 8 * b);
      fence_proxy_async();
      mbar_arrive(bRFull + 8 * b);
    }

    // ---- CTA-level reduction of the per-particle statistics and the gradient read-out -------
    mbar_wait(bGFull, 0);
    tc_fence_after();
    // every MMA has retired: X stage 0 is free to serve as reduction scratch [128][kNS+1]
    float* scratch = reinterpret_cast<float*>(gbase + kOffX);
    const int ncol = 1 + kP + 2;
    float* out = partial + (size_t)blockIdx.x * S * ncol;
    __shared__ float s_nlive[kEpiWarps];
    __shared__ double s_lgam[kEpiWarps];
    {
      const float nl = warp_sum(n_live);
      const double lg = warp_sum(lgam);
      if (lane == 0) { s_nlive[warp] = nl; s_lgam[warp] = lg; }
    }
    // pass 1: st0
#pragma unroll
    for (int s = 0; s < kNS; ++s) scratch[tid * (kNS + 1) + s] = st0[s];
    epi_bar_sync();
    float tot0 = 0.f;
    const float tot1 = 0.f;  // intercept gradient: intercepts are routed to the fp32 kernel for now
    if (tid < kNS) for (int rr = 0; rr < kTileM; ++rr) tot0 += scratch[rr * (kNS + 1) + tid];
    const float cnt = s_nlive[0] + s_nlive[1] + s_nlive[2] + s_nlive[3];
    const double lgsum = s_lgam[0] + s_lgam[1] + s_lgam[2] + s_lgam[3];
    if (tid < S) {
      const int s = tid;
      const ParticleParam pp = sPar[s];
      float lp, dicpt, dscale = 0.f;
      if (FAMILY == MNF_NORMAL) {
        const float inv = 1.0f / pp.scale, iv = inv * inv;
        lp = -0.5f * iv * tot0 - cnt * (logf(pp.scale) + kLogSqrt2Pi);
        dicpt = iv * tot1;
        dscale = (tot0 * iv * inv - cnt * inv) * pp.dscale;
      } else if (FAMILY == MNF_BERNOULLI_LOGITS) {
        lp = tot0;
        dicpt = tot1;
      } else {
        lp = tot0 - (float)lgsum;
        dicpt = tot1;
      }
      out[s * ncol + 0] = lp;
      out[s * ncol + 1 + kP] = dicpt;
      out[s * ncol + 2 + kP] = dscale;
    }
    // G: M = kP = 64 accumulator layout puts feature j on TMEM lane (j % 16) + 32 * (j / 16)
    // (cute tmem_frg_1sm, M_MMA == 64), i.e. lanes 0-15 of each epilogue warp's quadrant.
    {
      const int j = warp * 16 + lane;
#pragma unroll
      for (int ch = 0; ch < kNS / 16; ++ch) {
        uint32_t v[16];
        tc_ld16(tmem + lane_base + kColG + ch * 16, v);
        tc_wait_ld();
        if (lane < 16) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const int s = ch * 16 + c;
            if (s < S) {
              float g = __uint_as_float(v[c]);
              if (FAMILY == MNF_NORMAL) {
                const float inv = 1.0f / sPar[s].scale;
                g *= inv * inv;
              }
              out[s * ncol + 1 + j] = g;
            }
          }
        }
      }
    }
    if (bad_value) atomicOr(status, MNF_ST_BAD_VALUE);
    tc_fence_before();
  }

  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
  }
}

}  // namespace tc
}  // namespace mnf
