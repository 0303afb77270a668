// Normal likelihood with a dense linear predictor and a row-independent scale, p <= 64 (a multiple
// of 4; columns past p are zero-filled by the TMA unit and cost tensor time, not HBM traffic), with
// or without a row mask:
// the sweep reduced to DATA-ONLY Gram statistics (the move site_sweep.cuh makes for scalar links),
// expanded around the particle MEAN (a0, theta0) so that no large sums cancel:
//
//   r0_i = y_i - a0 - x_i.theta0,   alpha_s = a_s - a0,   delta_s = theta_s - theta0
//   sum_i (y_i - a_s - x_i.theta_s)^2 = R2 - 2 alpha R1 + n alpha^2 - 2 delta.(c - alpha sx) + delta' A delta
//   A = X'X [64 x 64],  c = X'r0,  sx = X'1,  R1 = sum r0,  R2 = sum r0^2,  n = rows
//
// so log p, d/dtheta = (c - alpha sx - A delta) / sigma^2, d/da and d/dsigma of EVERY particle follow
// from one pass over X that does no per-particle work (gram_finish_kernel evaluates the closed
// forms in fp64). The dominant term R2 is summed directly from the residuals; the Gram matrix only
// enters through the particles' spread delta' A delta. Replaces, like dense_tc.cuh: aten::mv +
// MvBackward + Normal.log_prob and its autograd twins (mininf/core.py:241, TORCH normal.py:87-103)
// for all S particles.
//
// Tensor work per 128-row tile: 16 tcgen05.mma kind::tf32 of M = N = 64, K = 8 whose A AND B
// operand are the SAME MN-major shared-memory image (A = X' needs the feature index contiguous,
// B = X too) - 2 N p^2 flop per step instead of 4 N p S, one TMA image per tile instead of two,
// no TMEM epilogue. X is rounded to TF32 (nearest even) by the TMA unit, products are exact,
// accumulation is fp32 in TMEM for kFlush tiles, then fp32 in registers per CTA, then fp64 across
// CTAs. r0, c, sx, R1, R2 are computed by the four SIMT warps from the same shared-memory image (the
// SWIZZLE_128B_ATOM_32B address map is applied by hand) in fp32 FMAs, summed in fp64.
//
// Warps: 0-3 residuals and X'r0 / X'1, 4 TMA producer, 6 y staging, 7 TMEM allocation + MMA issue,
// 8-11 drain the Gram accumulators (one TMEM lane quadrant each); 5 idles. Four accumulators rotate,
// so the MMA issuer, the SIMT warps and the drains are coupled only through the ring and never wait
// for each other in steady state.
#pragma once

#include "dense_tc.cuh"
#include "dense_tcr.cuh"

namespace mnf {
namespace gram {

using tc::elect_one;
using tc::idesc_tf32;
using tc::kAtomBytes;
using tc::kP;
using tc::kTileM;
using tc::kXImageBytes;
using tc::mbar_arrive;
using tc::mbar_arrive_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::smem_desc;
using tc::smem_u32;
using tc::tc_commit;
using tc::tc_fence_after;
using tc::tc_fence_before;
using tc::tc_ld32;
using tc::tc_wait_ld;
using tc::tma_load_2d;
using tcr::tc_mma_ss;

constexpr int kStages = 6;       // MN-major image ring (32 KB each)
constexpr int kFlush = 2;        // tiles accumulated in TMEM before the Gram tile is drained (truncating adds)
constexpr int kFlush64 = 64;     // tiles per fp32 run of the X'r0 / X'1 sums before they move to fp64
constexpr int kSimtWarps = 4;
constexpr int kWarpTma = 4, kWarpY = 6, kWarpMma = 7;
constexpr int kWarpDrain0 = 8;   // warps 8..11: TMEM lane quadrants 0..3
constexpr int kAcc = 4;          // Gram accumulators in tensor memory
constexpr int kThreads = 12 * 32;
constexpr uint32_t kYBytes = kTileM * 4;
constexpr uint32_t kTmemCols = 256;              // four 64-column accumulators

constexpr uint32_t kOffX = 0;
constexpr uint32_t kOffY = kOffX + kStages * kXImageBytes;
constexpr uint32_t kOffBar = kOffY + kStages * kYBytes;
constexpr uint32_t kNumBars = 3 * kStages + 2 * kAcc;         // full, empty, ready per stage; g_full, g_empty per accumulator
constexpr uint32_t kOffMisc = kOffBar + 8 * kNumBars;         // tmem slot
constexpr uint32_t kOffG = kOffMisc + 64;                     // this CTA's Gram rows [64][65] fp32 (written once)
constexpr uint32_t kOffVec = kOffG + kP * (kP + 1) * 4;       // [4 warps][2][64] doubles: X'r0, X'1
constexpr uint32_t kOffScal = kOffVec + kSimtWarps * 2 * kP * 8;   // [4 warps][2 halves][R1, R2] doubles, then n
constexpr uint32_t kOffCenter = kOffScal + (kSimtWarps * 4 + 1) * 8 + 8;   // theta0 [64] floats, a0
constexpr uint32_t kSmemBytes = kOffCenter + (kP + 4) * 4 + 1024 /* alignment slack */;
static_assert(kSmemBytes <= 227 * 1024, "shared-memory budget");

// per-CTA output: A [64][64], c [64], sx [64], R1, R2, n
constexpr int kCtaFloats = kP * kP + 2 * kP + 3;

// The expansion point: the particle mean of theta_j and of the intercept, summed in a fixed order in
// fp32 so that every CTA of the sweep and the finish kernel use bit-identical values.
__device__ __forceinline__ float center_theta(const mnf_dense_site_t& site, const float* __restrict__ z, int S, int D,
                                              int j) {
  if (j >= site.p) return 0.f;                         // p < 64: the TMA unit zero-fills those columns
  float t = 0.f;
#pragma unroll 8
  for (int s = 0; s < S; ++s) t += z[(int64_t)s * D + site.theta_lat + j];
  return t / (float)S;
}
__device__ __forceinline__ float center_icpt(const mnf_dense_site_t& site, const float* __restrict__ z, int S, int D) {
  float t = 0.f;
  if (site.icpt_lat >= 0)
    for (int s = 0; s < S; ++s) t += z[(int64_t)s * D + site.icpt_lat];
  return site.icpt_const + t / (float)S;
}

// MASKED: rows whose mask byte is zero contribute nothing. The y warp stages their response as NaN;
// the SIMT warps take that as the row's liveness, zero the row in the shared-memory image (and in
// their registers) and only then release the stage to the MMA issuer through a third barrier, so
// the Gram product sees X with the masked rows removed.
template <bool MASKED>
__global__ void __launch_bounds__(kThreads, 1)
dense_gram_kernel(const __grid_constant__ CUtensorMap map_mn, mnf_dense_site_t site, const float* __restrict__ z,
                  int S, int D, float* __restrict__ cta_out, uint32_t* __restrict__ status, uint32_t dev_skip) {
  // dev_skip (MNF_GRAM_DEV_SKIP, timing experiments only, results are wrong): 1 = no MMAs, 2 = no SIMT sums,
  // bits 8.. = tiles per TMEM accumulation run instead of kFlush
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzled images: 1024-byte alignment
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sX = base + kOffX, sY = base + kOffY;
  const uint32_t bars = base + kOffBar;
  const uint32_t bFull = bars, bEmpty = bFull + 8 * kStages;
  const uint32_t bGFull = bEmpty + 8 * kStages, bGEmpty = bGFull + 8 * kAcc;
  const uint32_t bReady = bGEmpty + 8 * kAcc;         // MASKED only: image cleaned by the SIMT warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + kOffMisc);
  float* sG = reinterpret_cast<float*>(gbase + kOffG);
  double* sVec = reinterpret_cast<double*>(gbase + kOffVec);
  double* sScal = reinterpret_cast<double*>(gbase + kOffScal);
  float* sCenter = reinterpret_cast<float*>(gbase + kOffCenter);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int flush = (dev_skip >> 8) != 0 ? (int)(dev_skip >> 8) : kFlush;   // developer override of kFlush
  const int64_t n_tiles = (site.n_rows + kTileM - 1) / kTileM;
  const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bFull + 8 * i, 2);                   // arrive.expect_tx of the producer + the y warp
      mbar_init(bEmpty + 8 * i, 1 + kSimtWarps);     // tcgen05.commit of the Gram product + the SIMT warps
      mbar_init(bReady + 8 * i, kSimtWarps);
    }
    for (int i = 0; i < kAcc; ++i) {
      mbar_init(bGFull + 8 * i, 1);
      mbar_init(bGEmpty + 8 * i, 4 * 32);            // the four drain warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kWarpMma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == kWarpTma && lane == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_mn) : "memory");
  for (int i = tid; i < kP * (kP + 1); i += kThreads) sG[i] = 0.f;
  if (tid < kP) sCenter[tid] = center_theta(site, z, S, D, tid);
  if (tid == kP) sCenter[kP] = center_icpt(site, z, S, D);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == kWarpTma) {
    // ================= TMA producer (one elected lane) ========================================
    if (elect_one()) {
      for (int64_t k = 0; k < my_tiles; ++k) {
        const int st = (int)(k % kStages);
        const int row0 = (int)((blockIdx.x + k * gridDim.x) * kTileM);
        mbar_wait(bEmpty + 8 * st, (uint32_t)(((k / kStages) & 1) ^ 1));
        mbar_arrive_expect_tx(bFull + 8 * st, kXImageBytes);
#pragma unroll
        for (int a = 0; a < kP / 32; ++a)
          tma_load_2d(sX + (uint32_t)st * kXImageBytes + a * kAtomBytes, &map_mn, a * 32, row0, bFull + 8 * st);
      }
    }
    __syncwarp();
  } else if (warp == kWarpY) {
    // ================= y warp: the tile's responses (0 past the end; NaN where masked), checks, n ===
    const bool y_vec = reinterpret_cast<uintptr_t>(site.y) % 16 == 0 &&
                       (!MASKED || reinterpret_cast<uintptr_t>(site.mask) % 4 == 0);
    int64_t cnt = 0;
    bool bad_value = false;
    auto fetch = [&](int64_t k, float4& v, uint32_t& m) {
      const int64_t row = (blockIdx.x + k * gridDim.x) * kTileM + lane * 4;
      if (y_vec && row + 4 <= site.n_rows) {
        v = __ldg(reinterpret_cast<const float4*>(site.y + row));
        m = MASKED ? __ldg(reinterpret_cast<const uint32_t*>(site.mask + row)) : 0x01010101u;
      } else {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        m = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (row + q < site.n_rows) {
            t[q] = __ldg(site.y + row + q);
            m |= ((MASKED ? (uint32_t)__ldg(site.mask + row + q) : 1u) != 0 ? 1u : 0u) << (8 * q);
          }
        v = make_float4(t[0], t[1], t[2], t[3]);
      }
    };
    // Responses of the next kStages tiles stay in registers: one outstanding load per tile would tie
    // the tile rate to the loaded HBM latency (~1 us), which is above the 0.75 us a tile may take.
    float4 yq[kStages];
    uint32_t mq[kStages];
#pragma unroll
    for (int i = 0; i < kStages; ++i) {
      yq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      mq[i] = 0;
      if (i < my_tiles) fetch(i, yq[i], mq[i]);
    }
    for (int64_t k0 = 0; k0 < my_tiles; k0 += kStages) {
#pragma unroll
      for (int i = 0; i < kStages; ++i) {
        const int64_t k = k0 + i;                      // k % kStages == i
        if (k < my_tiles) {
          float yv[4] = {yq[i].x, yq[i].y, yq[i].z, yq[i].w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool live = ((mq[i] >> (8 * q)) & 0xFFu) != 0;     // false past the end
            if (live) {
              ++cnt;
              if (yv[q] != yv[q]) bad_value = true;    // a live NaN response is outside the support
            } else if (MASKED) {
              yv[q] = __int_as_float(0x7fc00000);      // masked or past the end: not a row
            }
          }
          mbar_wait(bEmpty + 8 * i, (uint32_t)(((k / kStages) & 1) ^ 1));
          tc::sts128(sY + (uint32_t)i * kYBytes + lane * 16, __float_as_uint(yv[0]), __float_as_uint(yv[1]),
                     __float_as_uint(yv[2]), __float_as_uint(yv[3]));
          __syncwarp();
          if (lane == 0) mbar_arrive(bFull + 8 * i);
          if (k + kStages < my_tiles) fetch(k + kStages, yq[i], mq[i]);
        }
      }
    }
    if (bad_value) atomicOr(status, MNF_ST_BAD_VALUE);
    const double n_live = warp_sum((double)cnt);
    if (lane == 0) sScal[kSimtWarps * 4] = n_live;
  } else if (warp == kWarpMma) {
    // ================= MMA issuer: A and B are the same MN-major image =========================
    constexpr uint32_t idesc = idesc_tf32(kP, kP, 1, 1);            // M = N = 64, both MN-major
    const uint64_t dMN = smem_desc(sX, kAtomBytes, 512, 1);         // SWIZZLE_128B_BASE32B
    const uint32_t d_lo = (uint32_t)dMN, d_hi = (uint32_t)(dMN >> 32);
    for (int64_t k = 0; k < my_tiles; ++k) {
      const int st = (int)(k % kStages);
      const int64_t grp = k / flush;
      const uint32_t gb = (uint32_t)(grp % kAcc);
      const bool first = (k % flush) == 0;
      const bool last = (k % flush) == flush - 1 || k == my_tiles - 1;
      mbar_wait((MASKED ? bReady : bFull) + 8 * st, (uint32_t)((k / kStages) & 1));
      if (first) mbar_wait(bGEmpty + 8 * gb, (uint32_t)(((grp / kAcc) & 1) ^ 1));
      tc_fence_after();
      if (dev_skip & 1u) {
        if (lane == 0) {
          mbar_arrive(bEmpty + 8 * st);
          if (last) mbar_arrive(bGFull + 8 * gb);
        }
      } else if (elect_one()) {
        const uint32_t lo = d_lo + (uint32_t)st * (kXImageBytes >> 4);
        const uint32_t d = tmem + gb * kP;
#pragma unroll
        for (int ks = 0; ks < kTileM / 8; ++ks)
          tc_mma_ss(d, lo + ks * 64, d_hi, lo + ks * 64, d_hi, idesc, (!first || ks > 0) ? 1u : 0u);
        tc_commit(bEmpty + 8 * st);
        if (last) tc_commit(bGFull + 8 * gb);
      }
      __syncwarp();
    }
  } else if (warp < kSimtWarps) {
    // ================= SIMT warps: r0 = y - a0 - x.theta0, c = X'r0, sx = X'1, R1, R2 =============
    // Half-warp h = lane / 16 takes row 32 warp + 2 r + h of the tile, lane l' = lane % 16 its
    // features 4 l' .. 4 l' + 3: atom l' / 8, 16-byte chunk l' % 8 of the 128-byte row, moved by the
    // swizzle to 32-byte chunk ((l' % 8) / 2) ^ (row % 4).
    const int lq = lane & 15, half = lane >> 4;
    const uint32_t atom_off = (uint32_t)(lq >> 3) * kAtomBytes;
    const uint32_t c16 = (uint32_t)(lq & 7);
    const float4 t0 = *reinterpret_cast<const float4*>(sCenter + 4 * lq);
    const float a0 = sCenter[kP];
    const bool need_sx = site.icpt_lat >= 0;           // without a latent intercept every alpha_s is zero
    float by[4] = {0.f, 0.f, 0.f, 0.f}, bx[4] = {0.f, 0.f, 0.f, 0.f};
    double dby[4] = {0.0, 0.0, 0.0, 0.0}, dbx[4] = {0.0, 0.0, 0.0, 0.0};
    double dr1 = 0.0, dr2 = 0.0;
    for (int64_t k = 0; k < my_tiles; ++k) {
      const int st = (int)(k % kStages);
      mbar_wait(bFull + 8 * st, (uint32_t)((k / kStages) & 1));
      const uint32_t img = sX + (uint32_t)st * kXImageBytes + atom_off;
      const float* ys = reinterpret_cast<const float*>(gbase + kOffY + (size_t)st * kYBytes);
      const int64_t rows_left = site.n_rows - (blockIdx.x + k * gridDim.x) * kTileM;   // > 0
      float r1 = 0.f, r2 = 0.f;
      if (!(dev_skip & 2u)) {
        // pass A: this lane's four features of its half-warp's 16 rows, partial dot products with theta0
        float4 x[16];
        float v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const int row = warp * 32 + 2 * r + half;
          const uint32_t addr = img + (uint32_t)row * 128u + ((((c16 >> 1) ^ ((uint32_t)row & 3u)) << 5) | ((c16 & 1u) << 4));
          x[r] = tc::lds128(addr);
          v[r] = fmaf(x[r].x, t0.x, fmaf(x[r].y, t0.y, fmaf(x[r].z, t0.z, x[r].w * t0.w)));
        }
        // reduce-scatter butterfly over the 16 lanes of the half-warp: 15 shuffles instead of 64 (shuffles
        // and shared-memory loads share the pipe this kernel is short of); lane l' ends up with row 2 l' + h
#pragma unroll
        for (int w = 8; w >= 1; w >>= 1) {
          const bool up = (lq & w) != 0;
#pragma unroll
          for (int i = 0; i < w; ++i) {
            const float send = up ? v[i] : v[i + w];
            const float keep = up ? v[i + w] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, w);
          }
        }
        const int own = warp * 32 + 2 * lq + half;
        const float y_own = ys[own];
        const bool live_own = own < rows_left && (!MASKED || y_own == y_own);   // rows past the end: X and y are zeros
        const float res_own = live_own ? (y_own - a0) - v[0] : 0.f;
        r1 = res_own;
        r2 = res_own * res_own;
        if (MASKED) {
          // masked rows leave the image (and the registers) before the tensor core reads the stage
          const uint32_t live_bits = __ballot_sync(0xffffffffu, live_own);     // bit 16 h + r: row 2 r + h
#pragma unroll
          for (int r = 0; r < 16; ++r) {
            if (!((live_bits >> ((lane & 16) | r)) & 1u)) {
              const int row = warp * 32 + 2 * r + half;
              const uint32_t addr = img + (uint32_t)row * 128u + ((((c16 >> 1) ^ ((uint32_t)row & 3u)) << 5) | ((c16 & 1u) << 4));
              tc::sts128(addr, 0u, 0u, 0u, 0u);
              x[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores before the MMA's reads
          __syncwarp();
          if (lane == 0) mbar_arrive(bReady + 8 * st);
        }
        // pass B: c += x r0, sx += x with the row's residual fetched from its owner
#pragma unroll
        for (int r = 0; r < 16; ++r) {
          const float res = __shfl_sync(0xffffffffu, res_own, (lane & 16) | r);
          by[0] = fmaf(x[r].x, res, by[0]); by[1] = fmaf(x[r].y, res, by[1]);
          by[2] = fmaf(x[r].z, res, by[2]); by[3] = fmaf(x[r].w, res, by[3]);
        }
        if (need_sx) {                                 // X'1 only multiplies alpha_s = a_s - a0
#pragma unroll
          for (int r = 0; r < 16; ++r) { bx[0] += x[r].x; bx[1] += x[r].y; bx[2] += x[r].z; bx[3] += x[r].w; }
        }
      }
      if (MASKED && (dev_skip & 2u) && lane == 0) mbar_arrive(bReady + 8 * st);
      __syncwarp();
      if (lane == 0) mbar_arrive(bEmpty + 8 * st);
      dr1 += (double)r1;                               // one row per lane and tile: R2 is the dominant term
      dr2 += (double)r2;
      if ((k % kFlush64) == kFlush64 - 1) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          dby[c] += (double)by[c]; dbx[c] += (double)bx[c];
          by[c] = 0.f; bx[c] = 0.f;
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      dby[c] += (double)by[c]; dbx[c] += (double)bx[c];
      dby[c] += __shfl_xor_sync(0xffffffffu, dby[c], 16);
      dbx[c] += __shfl_xor_sync(0xffffffffu, dbx[c], 16);
    }
    if (half == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        sVec[(warp * 2 + 0) * kP + 4 * lq + c] = dby[c];
        sVec[(warp * 2 + 1) * kP + 4 * lq + c] = dbx[c];
      }
    }
    dr1 += __shfl_xor_sync(0xffffffffu, dr1, 8); dr2 += __shfl_xor_sync(0xffffffffu, dr2, 8);
    dr1 += __shfl_xor_sync(0xffffffffu, dr1, 4); dr2 += __shfl_xor_sync(0xffffffffu, dr2, 4);
    dr1 += __shfl_xor_sync(0xffffffffu, dr1, 2); dr2 += __shfl_xor_sync(0xffffffffu, dr2, 2);
    dr1 += __shfl_xor_sync(0xffffffffu, dr1, 1); dr2 += __shfl_xor_sync(0xffffffffu, dr2, 1);
    if (lq == 0) {                                     // one row per lane and tile: the half-warp's sums
      sScal[(warp * 2 + half) * 2 + 0] = dr1;
      sScal[(warp * 2 + half) * 2 + 1] = dr2;
    }
  } else if (warp >= kWarpDrain0) {
    // ================= drain warps: accumulator row m lives on TMEM lane (m % 16) + 32 (m / 16) ==
    // The running Gram row stays in REGISTERS (64 per lane): shared-memory bandwidth is what this
    // kernel runs out of first (TMA writes, MMA operand reads, the SIMT warps' loads and shuffles).
    const int q = warp - kWarpDrain0;                     // == warp % 4: this warp's lane quadrant
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    float g_acc[kP];
#pragma unroll
    for (int c = 0; c < kP; ++c) g_acc[c] = 0.f;
    const int64_t n_grp = (my_tiles + flush - 1) / flush;
    for (int64_t grp = 0; grp < n_grp; ++grp) {
      const uint32_t gb = (uint32_t)(grp % kAcc);
      mbar_wait(bGFull + 8 * gb, (uint32_t)((grp / kAcc) & 1));
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < kP / 32; ++ch) {
        uint32_t v[32];
        tc_ld32(tmem + lane_base + gb * kP + ch * 32, v);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; ++c) g_acc[ch * 32 + c] += __uint_as_float(v[c]);   // lanes >= 16 hold nothing
      }
      tc_fence_before();
      mbar_arrive(bGEmpty + 8 * gb);
    }
    tc_fence_before();
    if (lane < 16) {
      float* g_row = sG + (size_t)(q * 16 + lane) * (kP + 1);
#pragma unroll
      for (int c = 0; c < kP; ++c) g_row[c] = g_acc[c];
    }
  }

  __syncthreads();
  // ---- this CTA's statistics, fp32 (values up to ~1e6, summed in fp64 across CTAs) --------------
  float* out = cta_out + (size_t)blockIdx.x * kCtaFloats;
  for (int i = tid; i < kP * kP; i += kThreads) out[i] = sG[(i / kP) * (kP + 1) + (i % kP)];
  for (int i = tid; i < 2 * kP; i += kThreads) {
    const int which = i / kP, j = i % kP;
    double t = 0.0;
    for (int w = 0; w < kSimtWarps; ++w) t += sVec[(w * 2 + which) * kP + j];
    out[kP * kP + i] = (float)t;
  }
  if (tid < 2) {                                       // R1, R2: the eight half-warp sums in fixed order
    double t = 0.0;
    for (int w = 0; w < kSimtWarps * 2; ++w) t += sScal[w * 2 + tid];
    out[kP * kP + 2 * kP + tid] = (float)(my_tiles > 0 ? t : 0.0);
  }
  if (tid == 2) out[kP * kP + 2 * kP + 2] = (float)(my_tiles > 0 ? sScal[kSimtWarps * 4] : 0.0);
  if (warp == kWarpMma) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
  }
}

// totals over the CTAs in fixed order, fp64: [kCtaFloats]
__global__ void __launch_bounds__(256)
gram_reduce_kernel(const float* __restrict__ cta_out, int n_cta, double* __restrict__ total) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= kCtaFloats) return;
  double t = 0.0;
  for (int b = 0; b < n_cta; ++b) t += (double)cta_out[(size_t)b * kCtaFloats + i];
  total[i] = t;
}

// One block per particle, thread j = feature j: the closed forms around the particle mean, written
// as one row block [S][1 + p + 2] in the layout of the other dense kernels (log-density, d/dtheta,
// d/dintercept, d/d(scale link)) so the common reduction maps it into the step accumulator.
__global__ void __launch_bounds__(kP)
gram_finish_kernel(mnf_dense_site_t site, const double* __restrict__ total, const float* __restrict__ z, int S, int D,
                   float* __restrict__ rows, uint32_t* __restrict__ status) {
  __shared__ double s_delta[kP];
  const int s = blockIdx.x, j = threadIdx.x;
  const float* zs = z + (int64_t)s * D;
  const DenseParticle pp = dense_particle(site, zs);
  const int p = site.p;                              // <= 64; columns past p hold zeros everywhere
  s_delta[j] = j < p ? (double)zs[site.theta_lat + j] - (double)center_theta(site, z, S, D, j) : 0.0;
  __syncthreads();
  const double* A = total;
  const double* c = total + kP * kP;
  const double* sx = c + kP;
  const double R1 = total[kP * kP + 2 * kP], R2 = total[kP * kP + 2 * kP + 1], n = total[kP * kP + 2 * kP + 2];
  double ad = 0.0;                                   // (A delta)_j, A symmetric: column reads coalesce
  for (int k = 0; k < kP; ++k) ad = fma(A[k * kP + j], s_delta[k], ad);
  const double alpha = (double)pp.icpt - (double)center_icpt(site, z, S, D);
  const double cj = c[j] - alpha * sx[j];            // sum_i x_ij (r0_i - alpha)
  const double rj = cj - ad;                         // sum_i x_ij (r0_i - alpha - x_i.delta)
  double q1 = s_delta[j] * cj, q2 = s_delta[j] * rj, q3 = s_delta[j] * sx[j];
  q1 = warp_sum(q1); q2 = warp_sum(q2); q3 = warp_sum(q3);
  __shared__ double s_q[3][kP / 32];
  if ((j & 31) == 0) { s_q[0][j >> 5] = q1; s_q[1][j >> 5] = q2; s_q[2][j >> 5] = q3; }
  __syncthreads();
  const double t1 = s_q[0][0] + s_q[0][1], t2 = s_q[1][0] + s_q[1][1], t3 = s_q[2][0] + s_q[2][1];
  const double sigma = (double)pp.scale, inv = 1.0 / sigma, iv = inv * inv;
  const int ncol = 1 + p + 2;
  float* out = rows + (size_t)s * ncol;
  if (j < p) out[1 + j] = (float)(rj * iv);
  if (j == 0) {
    const double Q = R2 - 2.0 * alpha * R1 + n * alpha * alpha - t1 - t2;   // sum of squared residuals
    const double r1 = R1 - n * alpha - t3;                                  // sum of residuals
    out[0] = (float)(-0.5 * iv * Q - n * (log(sigma) + 0.91893853320467274178));
    out[1 + p] = (float)(r1 * iv);
    out[2 + p] = (float)((Q * iv * inv - n * inv) * (double)pp.dscale);
    if (!(pp.scale > 0.0f)) atomicOr(status, MNF_ST_BAD_PARAM);
  }
}

}  // namespace gram
}  // namespace mnf
