// Dense-link sweep, tcgen05 variant with observations on the TMEM lanes ("rows-on-lanes"):
// the tensor-core kernel for wide design matrices (p <= 64 * C features, p % 4 == 0, any C that
// fits; the TMA unit zero-fills the padding columns of the last chunk), at most 32 particles per
// sweep, with or without an intercept. Config C3 (minibatch logistic regression, p = 256,
// S = 16) runs here; dense_tc.cuh keeps the p = 64 / S <= 64 shape of config C2.
//
// One persistent CTA per SM walks 128-row tiles; a tile is streamed as C chunks of 64 features
// (32 KB per operand image):
//
//   eta[128 x 2NS]  = sum_c Xc[128 x 64] . Theta_c^T[64 x 2NS]  tcgen05.mma kind::tf32 M=128 N=2NS
//                     A = X chunk, K-major SWIZZLE_128B image; B = Theta, K-major, staged once as
//                     NS hi rows tf32(theta) and NS lo rows tf32(theta - hi): eta = hi column + lo
//                     column carries theta to 22 bits (a TF32-rounded Theta is a systematic
//                     per-particle error that does not average out over rows); + 8 cycles per MMA
//   R  [128 x NS]   = score(y, eta + intercept)                 epilogue: thread == row; written to
//                     shared memory as the K-major B operand of the gradient product
//   Gc [64 x NS]   += Xc^T[64 x 128] . R[128 x NS]              tcgen05.mma kind::tf32 M=64 N=NS
//                     A = X chunk, MN-major SWIZZLE_128B_BASE32B image; B = R
//
// Why this orientation: with particles on the MMA N dimension one MMA costs about
// N/2 + 11 + M/4 cycles (tools/umma_time.cu), i.e. 8*59 + 16*35 = 1032 cycles per chunk at
// NS = 16 against ~1100 cycles of HBM time per 32 KB chunk and SM, and tensor memory holds only
// (4 + 2C) * NS columns. The transposed orientation of dense_tc.cuh would need 256 columns for
// Theta alone at p = 256 and 75 + 43 cycles per MMA pair.
//
// Data movement: as in dense_tc.cuh both operand images of a chunk come from TMA with the
// TFLOAT32 tensor-map type (round to nearest even in flight, rows past the end zero-filled). The
// K-major image of tile i+1 and the MN-major image of tile i are streamed through two small
// rings; HBM latency is hidden by L2 prefetches (cp.async.bulk.prefetch.tensor) issued
// kPrefetchChunks ahead of the K-ring loads, so both rings hit L2 and X leaves HBM once. The MMA
// warp interleaves, chunk by chunk, the eta product of tile t with the gradient product of tile
// t - 1: both rings drain at a steady rate and need only cover the L2 latency.
//
// Warps (256 threads): 0-3 epilogue (thread = tile row = TMEM lane), 4 MN-ring TMA, 5 K-ring TMA
// and L2 prefetch, 6 stages y (and the mask) of each tile in shared memory, 7 TMEM allocation and
// every MMA.
//
// Precision mode: the same as dense_tc.cuh (X rounded to TF32 by the TMA unit, Theta as hi + lo
// TF32 pairs, scores rounded to TF32, fp32 accumulate in TMEM, the gradient accumulators
// ping-ponged and drained every kFlush tiles, fp32/fp64 SIMT for log-densities and sums).
//
// Replaces: aten::mv / addmv_ and MvBackward of `X @ theta (+ intercept)`
// (examples/minibatch.md:33, tests/test_mininf.py:11) and the Normal / Bernoulli / Poisson
// log_prob chains with their autograd twins (mininf/core.py:241), for all particles in one pass.
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "dense_simt.cuh"
#include "dense_tc.cuh"

namespace mnf {
namespace tcr {

using tc::elect_one;
using tc::fence_proxy_async;
using tc::idesc_tf32;
using tc::mbar_arrive;
using tc::mbar_arrive_expect_tx;
using tc::mbar_init;
using tc::mbar_wait;
using tc::rn_tf32;
using tc::smem_desc;
using tc::smem_u32;
using tc::tc_commit;
using tc::tc_fence_after;
using tc::tc_fence_before;
using tc::tc_wait_ld;
using tc::tma_load_2d;

constexpr int kTileM = 128;            // rows per tile: MMA M of the eta product, K of the gradient product
constexpr int kChunk = 64;             // features per operand chunk
constexpr int kMaxStages = 4;          // ring depths (chunks) are chosen by the host, 2..4 each
// The gradient product of a tile is issued one tile after its eta product. (Two tiles would take
// the epilogue off the critical path entirely but measured 25 % slower at p = 256: the MN-major
// re-read of a tile then arrives too long after its first touch and misses L2.)
constexpr int kFlush = 8;              // tiles accumulated in TMEM before the gradient tiles are drained
constexpr int kStatFlush = 256;        // tiles between fp32 -> fp64 hand-overs of the row statistics
#ifndef MNF_TCR_PREFETCH
#define MNF_TCR_PREFETCH 4
#endif
constexpr int kPrefetchChunks = MNF_TCR_PREFETCH;   // L2 prefetch distance ahead of the K-ring loads
constexpr int kEpiWarps = 4;
constexpr int kWarpMnTma = 4, kWarpKTma = 5, kWarpY = 6, kMmaWarp = 7;
constexpr int kThreads = 8 * 32;
constexpr uint32_t kTmemCols = 512;

constexpr uint32_t kAtomBytes = kTileM * 128;          // 128 rows x 32 fp32
constexpr uint32_t kChunkBytes = 2 * kAtomBytes;       // 32 KB per image of a chunk
constexpr uint32_t kBarBytes = 8 * (4 * kMaxStages + 8 + 2);   // + y_full, y_empty
constexpr uint32_t kYBytes = kTileM * 4;               // one tile of responses (masked-out rows are NaN)

// dynamic shared memory map (bytes from the 1024-aligned base)
struct Layout {
  uint32_t off_k, off_mn, off_theta, off_r, off_bar, off_misc, off_par, off_stat, off_y, total;
};
__host__ __device__ inline Layout make_layout(int NS, int C, int k_stages, int mn_stages) {
  Layout l;
  l.off_k = 0;
  l.off_mn = l.off_k + (uint32_t)k_stages * kChunkBytes;
  l.off_theta = l.off_mn + (uint32_t)mn_stages * kChunkBytes;
  l.off_r = l.off_theta + (uint32_t)(2 * C) * (uint32_t)(2 * NS) * 128u;   // Theta: 2C atoms of 2NS (hi, lo) rows x 128 B
  l.off_bar = l.off_r + 2u * 4u * (uint32_t)NS * 128u;                // R: two buffers of four atoms
  l.off_misc = l.off_bar + kBarBytes;                                 // tmem slot + two fp64 counters
  l.off_par = l.off_misc + 32;                                        // DenseParticle[NS], 16 B apart
  l.off_stat = l.off_par + (uint32_t)NS * 16u;                        // double [NS][2]
  l.off_y = l.off_stat + (uint32_t)NS * 16u;
  l.total = l.off_y + kYBytes + 1024u /* alignment slack */;
  return l;
}

// D[tmem] (+)= A[smem descriptor] . B[smem descriptor]
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                          uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* map, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1)
               : "memory");
}
template <int NS>
__device__ __forceinline__ void tc_ld_row(uint32_t taddr, uint32_t (&v)[NS]) {
  if constexpr (NS == 16) tc::tc_ld16(taddr, v);
  else tc::tc_ld32(taddr, v);
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// byte offset of element (row n, k) of a K-major SWIZZLE_128B operand with NS rows: atoms of 32
// k-elements, 128-byte rows, 16-byte chunks XOR-ed with the row index modulo 8
__device__ __forceinline__ uint32_t kmajor_offset(int NS, int n, int k) {
  return (uint32_t)(k >> 5) * (uint32_t)(NS * 128) + (uint32_t)n * 128u +
         ((((uint32_t)(k & 31) >> 2) ^ ((uint32_t)n & 7u)) << 4) + ((uint32_t)k & 3u) * 4u;
}

// partial layout per CTA: [S][ncol], ncol = 1 + p + 2 (same as the other dense kernels)
template <int FAMILY, int NS, bool ICPT>
__global__ void __launch_bounds__(kThreads, 1)
dense_tcr_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_mn,
                 mnf_dense_site_t site, const float* __restrict__ z, int S, int D, int C, int k_stages,
                 int mn_stages, float* __restrict__ partial, uint32_t* __restrict__ status) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // swizzled images need 1024-byte alignment
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const Layout L = make_layout(NS, C, k_stages, mn_stages);
  const uint32_t sK = base + L.off_k, sMN = base + L.off_mn, sTheta = base + L.off_theta, sR = base + L.off_r;
  const uint32_t bars = base + L.off_bar;
  const uint32_t bKFull = bars, bKEmpty = bKFull + 8 * kMaxStages;
  const uint32_t bMnFull = bKEmpty + 8 * kMaxStages, bMnEmpty = bMnFull + 8 * kMaxStages;
  const uint32_t bEtaFull = bMnEmpty + 8 * kMaxStages, bRReady = bEtaFull + 16;
  const uint32_t bGFull = bRReady + 16, bGEmpty = bGFull + 16;
  const uint32_t bYFull = bGEmpty + 16, bYEmpty = bYFull + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + L.off_misc);
  double* counters = reinterpret_cast<double*>(gbase + L.off_misc + 16);   // [0] live rows, [1] sum lgamma(y+1)
  auto par_at = [&](int n) { return reinterpret_cast<DenseParticle*>(gbase + L.off_par + 16 * n); };
  double* sStat = reinterpret_cast<double*>(gbase + L.off_stat);           // [NS][2]: statistic, sum of scores

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int p = site.p;                 // features; C * kChunk >= p, the padding columns are zero
  const int p_pad = C * kChunk;
  const int ncol = 1 + p + 2;

  const int64_t n_tiles = (site.n_rows + kTileM - 1) / kTileM;
  const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  // ---- one-time setup ----------------------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < kMaxStages; ++i) {
      mbar_init(bKFull + 8 * i, 1);
      mbar_init(bKEmpty + 8 * i, 1);
      mbar_init(bMnFull + 8 * i, 1);
      mbar_init(bMnEmpty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bEtaFull + 8 * i, 1);
      mbar_init(bRReady + 8 * i, kEpiWarps * 32);
      mbar_init(bGFull + 8 * i, 1);
      mbar_init(bGEmpty + 8 * i, kEpiWarps * 32);
    }
    mbar_init(bYFull, 1);
    mbar_init(bYEmpty, kEpiWarps * 32);
    counters[0] = 0.0;
    counters[1] = 0.0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == kWarpMnTma && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_mn) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
  }
  for (int n = tid; n < NS; n += kThreads) {
    DenseParticle dp;
    dp.icpt = 0.f; dp.scale = 1.f; dp.dscale = 0.f;
    if (n < S) {
      dp = dense_particle(site, z + (int64_t)n * D);
      if (FAMILY == MNF_NORMAL && !(dp.scale > 0.0f)) atomicOr(status, MNF_ST_BAD_PARAM);
    }
    *par_at(n) = dp;
    sStat[2 * n] = 0.0;
    sStat[2 * n + 1] = 0.0;
  }
  // Theta as the B operand of the eta product: particle n is row n (hi half) and row NS + n (lo
  // half), feature j is k; spare particle rows are zero. Consecutive threads take consecutive
  // features of one particle.
  for (int i = tid; i < NS * p_pad; i += kThreads) {
    const int n = i / p_pad, j = i - n * p_pad;
    const float th = (n < S && j < p) ? z[(int64_t)n * D + site.theta_lat + j] : 0.f;
    const uint32_t hi = rn_tf32(th);
    sts32(sTheta + kmajor_offset(2 * NS, n, j), hi);
    sts32(sTheta + kmajor_offset(2 * NS, NS + n, j), rn_tf32(th - __uint_as_float(hi)));
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t col_eta = 0;                   // + b * 2NS (hi columns, then lo columns)
  const uint32_t col_g = 4 * NS;                // + (gb * C + c) * NS

  if (warp == kWarpMnTma) {
    // ================= MN-major ring: TMA producer (one elected lane) =========================
    if (elect_one()) {
      int st = 0;
      uint32_t ph = 0;
      for (int64_t k = 0; k < my_tiles; ++k) {
        const int row0 = (int)((blockIdx.x + k * gridDim.x) * kTileM);
        for (int c = 0; c < C; ++c) {
          mbar_wait(bMnEmpty + 8 * st, ph ^ 1u);
          mbar_arrive_expect_tx(bMnFull + 8 * st, kChunkBytes);
          tma_load_2d(sMN + (uint32_t)st * kChunkBytes, &map_mn, c * kChunk, row0, bMnFull + 8 * st);
          tma_load_2d(sMN + (uint32_t)st * kChunkBytes + kAtomBytes, &map_mn, c * kChunk + 32, row0, bMnFull + 8 * st);
          if (++st == mn_stages) { st = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == kWarpKTma) {
    // ================= K-major ring: TMA producer + L2 prefetch (one elected lane) ============
    if (elect_one()) {
      const int64_t total = my_tiles * C;
      auto prefetch = [&](int64_t q) {
        if (q < total) {
          const int64_t k = q / C;
          const int c = (int)(q - k * C);
          const int row0 = (int)((blockIdx.x + k * gridDim.x) * kTileM);
          tma_prefetch_2d(&map_k, c * kChunk, row0);
          tma_prefetch_2d(&map_k, c * kChunk + 32, row0);
        }
      };
      for (int q = 0; q < kPrefetchChunks; ++q) prefetch(q);
      int st = 0;
      uint32_t ph = 0;
      int64_t q = 0;
      for (int64_t k = 0; k < my_tiles; ++k) {
        const int row0 = (int)((blockIdx.x + k * gridDim.x) * kTileM);
        for (int c = 0; c < C; ++c, ++q) {
          prefetch(q + kPrefetchChunks);
          mbar_wait(bKEmpty + 8 * st, ph ^ 1u);
          mbar_arrive_expect_tx(bKFull + 8 * st, kChunkBytes);
          tma_load_2d(sK + (uint32_t)st * kChunkBytes, &map_k, c * kChunk, row0, bKFull + 8 * st);
          tma_load_2d(sK + (uint32_t)st * kChunkBytes + kAtomBytes, &map_k, c * kChunk + 32, row0, bKFull + 8 * st);
          if (++st == k_stages) { st = 0; ph ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == kWarpY) {
    // ================= y warp: responses of each tile -> shared memory =========================
    // The epilogue threads must not have global loads of their own in flight: their
    // fence.proxy.async (R tile -> tensor core) waits for them, which would put a memory latency
    // on the per-tile critical path. This warp loads y and the mask four rows per lane, two
    // tiles ahead in registers, folds the mask into the value (NaN = masked out or past the end),
    // and keeps the particle-independent sums (live rows, sum log y!).
    float* sY = reinterpret_cast<float*>(gbase + L.off_y);
    constexpr int kYDepth = 4;     // tiles in flight: one tile time is shorter than a memory latency
    float yq[kYDepth][4];
    auto fetch = [&](int64_t k, float (&out)[4]) {
      const int64_t row = (blockIdx.x + k * gridDim.x) * kTileM + lane * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        out[q] = __int_as_float(0x7fc00000);
        if (k < my_tiles && row + q < site.n_rows && (site.mask == nullptr || __ldg(site.mask + row + q) != 0)) {
          const float y = __ldg(site.y + row + q);
          out[q] = y;
          if (y != y) out[q] = __int_as_float(0x7fc00001);   // a live NaN: reported below
        }
      }
    };
#pragma unroll
    for (int i = 0; i < kYDepth; ++i) fetch(i, yq[i]);
    int live_total = 0;            // per-lane sums, combined once after the last tile
    double lgam_total = 0.0;
    bool bad_value = false;
    for (int64_t k0 = 0; k0 < my_tiles; k0 += kYDepth) {
#pragma unroll
      for (int i = 0; i < kYDepth; ++i) {
        const int64_t k = k0 + i;
        if (k < my_tiles) {
          mbar_wait(bYEmpty, (uint32_t)((k & 1) ^ 1));
          *reinterpret_cast<float4*>(sY + lane * 4) = make_float4(yq[i][0], yq[i][1], yq[i][2], yq[i][3]);
          __syncwarp();
          if (lane == 0) mbar_arrive(bYFull);
          float lgam = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float y = yq[i][q];
            if (__float_as_int(y) == 0x7fc00001) bad_value = true;
            if (y == y) {
              ++live_total;
              if (!in_support(FAMILY, y)) bad_value = true;
              if (FAMILY == MNF_POISSON) lgam += log_factorial(y);
            }
          }
          if (FAMILY == MNF_POISSON) lgam_total += (double)lgam;
          fetch(k + kYDepth, yq[i]);
        }
      }
    }
    const double live_rows = warp_sum((double)live_total);
    if (FAMILY == MNF_POISSON) lgam_total = warp_sum(lgam_total);
    if (lane == 0) {
      counters[0] = live_rows;
      counters[1] = lgam_total;
    }
    if (bad_value) atomicOr(status, MNF_ST_BAD_VALUE);
    asm volatile("bar.sync 1, %0;" ::"n"((kEpiWarps + 1) * 32) : "memory");
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer: warp-uniform loops, one elected lane issues ================
    constexpr uint32_t idesc_eta = idesc_tf32(kTileM, 2 * NS, 0, 0);   // M=128 N=2NS, A and B K-major
    constexpr uint32_t idesc_g = idesc_tf32(kChunk, NS, 1, 0);     // M=64  N=NS, A MN-major, B K-major
    const uint64_t dK = smem_desc(sK, 16, 1024, 2);                // K-major SWIZZLE_128B
    const uint64_t dMN = smem_desc(sMN, kAtomBytes, 512, 1);       // MN-major SWIZZLE_128B_BASE32B
    const uint64_t dT = smem_desc(sTheta, 16, 1024, 2);
    const uint64_t dR = smem_desc(sR, 16, 1024, 2);
    const uint32_t dK_lo = (uint32_t)dK, dK_hi = (uint32_t)(dK >> 32);
    const uint32_t dMN_lo = (uint32_t)dMN, dMN_hi = (uint32_t)(dMN >> 32);
    const uint32_t dT_lo = (uint32_t)dT, dT_hi = (uint32_t)(dT >> 32);
    const uint32_t dR_lo = (uint32_t)dR, dR_hi = (uint32_t)(dR >> 32);
    constexpr uint32_t kBAtom16 = (uint32_t)(NS * 128) >> 4;        // R atom stride (NS rows), 16-byte units
    constexpr uint32_t kTAtom16 = (uint32_t)(2 * NS * 128) >> 4;    // Theta atom stride (2NS rows)
    int kst = 0, mst = 0;
    uint32_t kph = 0, mph = 0;
    // Step t interleaves, chunk by chunk, the eta product of tile t with the gradient product of
    // tile t - 1, so both rings drain at a steady rate and the epilogue of a tile overlaps
    // tensor-core work of its neighbours.
    for (int64_t t = 0; t <= my_tiles; ++t) {
      const bool has_eta = t < my_tiles;
      const int64_t kk = t - 1;
      const bool has_g = kk >= 0;
      const uint32_t be = (uint32_t)(t & 1), bg = (uint32_t)(kk & 1);
      const int64_t grp = has_g ? kk / kFlush : 0;
      const uint32_t gb = (uint32_t)(grp & 1);
      const bool first = has_g && (kk % kFlush) == 0;
      const bool last = has_g && ((kk % kFlush) == kFlush - 1 || kk == my_tiles - 1);
      for (int c = 0; c < C; ++c) {
        if (has_eta) {
          // eta buffer `be` was read by the epilogue of tile t-2, whose r_ready was awaited in the
          // previous step
          mbar_wait(bKFull + 8 * kst, kph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t d = tmem + col_eta + be * (2 * NS);
            const uint32_t a_lo = dK_lo + (uint32_t)kst * (kChunkBytes >> 4);
            const uint32_t b_lo = dT_lo + (uint32_t)(2 * c) * kTAtom16;
#pragma unroll
            for (int a = 0; a < 2; ++a) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                tc_mma_ss(d, a_lo + ((a * kAtomBytes + ks * 32) >> 4), dK_hi, b_lo + a * kTAtom16 + ks * 2, dT_hi,
                          idesc_eta, (c | a | ks) != 0 ? 1u : 0u);
              }
            }
            tc_commit(bKEmpty + 8 * kst);
            // the epilogue may start now: it writes the R buffer that the gradient product of
            // this step does not read
            if (c == C - 1) tc_commit(bEtaFull + 8 * be);
          }
          __syncwarp();
          if (++kst == k_stages) { kst = 0; kph ^= 1u; }
        }
        if (has_g) {
          if (c == 0) {
            mbar_wait(bRReady + 8 * bg, (uint32_t)((kk >> 1) & 1));
            if (first) mbar_wait(bGEmpty + 8 * gb, (uint32_t)(((grp >> 1) & 1) ^ 1));
          }
          mbar_wait(bMnFull + 8 * mst, mph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = dMN_lo + (uint32_t)mst * (kChunkBytes >> 4);
            const uint32_t b_lo = dR_lo + bg * (4 * kBAtom16);
            const uint32_t d = tmem + col_g + (gb * (uint32_t)C + (uint32_t)c) * NS;
#pragma unroll
            for (int ks = 0; ks < kTileM / 8; ++ks) {
              tc_mma_ss(d, a_lo + ks * 64, dMN_hi, b_lo + (ks >> 2) * kBAtom16 + (ks & 3) * 2, dR_hi, idesc_g,
                        (!first || ks > 0) ? 1u : 0u);
            }
            tc_commit(bMnEmpty + 8 * mst);
            if (last && c == C - 1) tc_commit(bGFull + 8 * gb);
          }
          __syncwarp();
          if (++mst == mn_stages) { mst = 0; mph ^= 1u; }
        }
      }
    }
  } else if (warp < kEpiWarps) {
    // ================= epilogue warps: thread == tile row == TMEM lane =========================
    const int trow = warp * 32 + lane;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    // gradient drain: lanes 0-15 of quadrant `warp` carry features 16*warp + lane of every chunk
    // (features past p exist only as zero padding of the last chunk)
    const int f_lane = 16 * warp + lane;
    auto owns = [&](int c) { return lane < 16 && kChunk * c + f_lane < p; };
    float* g_out = partial + (size_t)blockIdx.x * S * ncol + 1 + 16 * warp + lane;   // + n * ncol + 64 * c
    for (int c = 0; c < C; ++c)
      if (owns(c))
        for (int n = 0; n < S; ++n) g_out[(size_t)n * ncol + kChunk * c] = 0.f;
    int64_t n_drained = 0;
    auto drain = [&]() {
      const int64_t grp = n_drained;
      const uint32_t gb = (uint32_t)(grp & 1);
      mbar_wait(bGFull + 8 * gb, (uint32_t)((grp >> 1) & 1));
      tc_fence_after();
      for (int c = 0; c < C; ++c) {
        uint32_t v[NS];
        float old[NS];
        // all loads of the read-modify-write first: one L2 round trip per chunk instead of NS
        const bool g_owner = owns(c);
        if (g_owner) {
#pragma unroll
          for (int n = 0; n < NS; ++n) old[n] = n < S ? __ldcg(g_out + (size_t)n * ncol + kChunk * c) : 0.f;
        }
        tc_ld_row<NS>(tmem + lane_base + col_g + (gb * (uint32_t)C + (uint32_t)c) * NS, v);
        tc_wait_ld();
        if (g_owner) {
#pragma unroll
          for (int n = 0; n < NS; ++n)
            if (n < S) __stcg(g_out + (size_t)n * ncol + kChunk * c, old[n] + __uint_as_float(v[n]));
        }
      }
      tc_fence_before();
      mbar_arrive(bGEmpty + 8 * gb);
      ++n_drained;
    };

    float st[NS], sr[ICPT ? NS : 1];
#pragma unroll
    for (int n = 0; n < NS; ++n) st[n] = 0.f;
#pragma unroll
    for (int n = 0; n < (ICPT ? NS : 1); ++n) sr[n] = 0.f;
    float icpt[ICPT ? NS : 1];
    if (ICPT) {
#pragma unroll
      for (int n = 0; n < (ICPT ? NS : 1); ++n) icpt[n] = par_at(n)->icpt;
    }
    auto hand_over = [&]() {   // fp32 per-thread sums -> fp64 per-particle sums in shared memory
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        const float a = warp_sum(st[n]);
        st[n] = 0.f;
        float b = 0.f;
        if (ICPT) { b = warp_sum(sr[ICPT ? n : 0]); sr[ICPT ? n : 0] = 0.f; }
        if (lane == 0) {
          atomicAdd(&sStat[2 * n], (double)a);
          if (ICPT) atomicAdd(&sStat[2 * n + 1], (double)b);
        }
      }
    };

    const float* sY = reinterpret_cast<const float*>(gbase + L.off_y);

    for (int64_t k = 0; k < my_tiles; ++k) {
      const uint32_t b = (uint32_t)(k & 1);
      // this row's response, staged by the y warp (NaN = masked out or past the end)
      mbar_wait(bYFull, (uint32_t)(k & 1));
      const float y_raw = sY[trow];
      mbar_arrive(bYEmpty);
      const float live = y_raw == y_raw ? 1.f : 0.f;
      const float y = y_raw == y_raw ? y_raw : 0.f;
      // the gradient group that ended two tiles ago has been issued (its R tile was handed over
      // two iterations back), so waiting for its commit cannot deadlock
      if (k >= 2 && ((k - 2) % kFlush) == kFlush - 1) drain();
      if (k > 0 && (k % kStatFlush) == 0) hand_over();
      mbar_wait(bEtaFull + 8 * b, (uint32_t)((k >> 1) & 1));
      tc_fence_after();
      uint32_t v[NS], vlo[NS];
      tc_ld_row<NS>(tmem + lane_base + col_eta + b * (2 * NS), v);
      tc_ld_row<NS>(tmem + lane_base + col_eta + b * (2 * NS) + NS, vlo);
      tc_wait_ld();
      // R buffer b was last read by the gradient product of tile k-2, which the commit behind
      // eta_full of this tile covers (tcgen05.commit tracks every earlier MMA of the issuer)
      const uint32_t r_row = sR + b * (uint32_t)(4 * NS * 128) + (uint32_t)warp * (uint32_t)(NS * 128) +
                             ((uint32_t)lane & 3u) * 4u;
      const uint32_t chunk = (uint32_t)lane >> 2;
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        float eta = __uint_as_float(v[n]) + __uint_as_float(vlo[n]);
        if (ICPT) eta += icpt[ICPT ? n : 0];
        float score;
        if (FAMILY == MNF_NORMAL) {
          score = live * (y - eta);             // 1/sigma^2 applied at the end
          st[n] = fmaf(score, score, st[n]);
        } else if (FAMILY == MNF_BERNOULLI_LOGITS) {
          const float e = __expf(-fabsf(eta));
          const float inv = __fdividef(1.0f, 1.0f + e);
          const float sig = eta >= 0.f ? inv : e * inv;
          score = live * (y - sig);
          st[n] += live * (y * eta - (fmaxf(eta, 0.f) + __logf(1.0f + e)));   // softplus, abs. error ~1e-7
        } else {
          const float rate = __expf(eta);
          score = live * (y - rate);
          st[n] += live * fmaf(y, eta, -rate);
        }
        if (ICPT) sr[ICPT ? n : 0] += score;
        sts32(r_row + (uint32_t)n * 128u + ((chunk ^ ((uint32_t)n & 7u)) << 4), rn_tf32(score));
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(bRReady + 8 * b);
    }
    {
      const int64_t n_grp = (my_tiles + kFlush - 1) / kFlush;
      while (n_drained < n_grp) drain();
    }
    hand_over();
    // the y warp's counters are complete once it has passed the same barrier
    asm volatile("bar.sync 1, %0;" ::"n"((kEpiWarps + 1) * 32) : "memory");

    // ---- per-particle results -----------------------------------------------------------------
    if (FAMILY == MNF_NORMAL) {
      for (int n = 0; n < S; ++n) {
        const float sc = par_at(n)->scale;
        const float iv = 1.0f / (sc * sc);
        for (int c = 0; c < C; ++c)
          if (owns(c)) g_out[(size_t)n * ncol + kChunk * c] *= iv;
      }
    }
    if (trow < S) {
      const int n = trow;
      float* out = partial + ((size_t)blockIdx.x * S + n) * ncol;
      const DenseParticle pp = *par_at(n);
      const float cnt = (float)counters[0];
      const float st0 = (float)sStat[2 * n], sr0 = (float)sStat[2 * n + 1];
      float lp, gscale = 1.0f, dscale = 0.f;
      if (FAMILY == MNF_NORMAL) {
        const float inv = 1.0f / pp.scale, iv = inv * inv;
        lp = -0.5f * iv * st0 - cnt * (logf(pp.scale) + kLogSqrt2Pi);
        dscale = (st0 * iv * inv - cnt * inv) * pp.dscale;
        gscale = iv;
      } else if (FAMILY == MNF_BERNOULLI_LOGITS) {
        lp = st0;
      } else {
        lp = st0 - (float)counters[1];
      }
      out[0] = lp;
      out[1 + p] = ICPT ? sr0 * gscale : 0.f;
      out[2 + p] = dscale;
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols));
  }
}

}  // namespace tcr
}  // namespace mnf
