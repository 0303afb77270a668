// Dense-link sweep, tcgen05 variant (MNF_DENSE_TF32) for sm_100a.
//
// One persistent CTA per SM streams 128-row tiles of X from HBM exactly once and runs BOTH matrix
// products of the ELBO step on the 5th-generation tensor cores, in the transposed orientation
// that keeps every intermediate in tensor memory:
//
//   eta^T[2S x 128] = Theta[2S x p] . Xtile^T[p x 128]    tcgen05.mma kind::tf32, M=128 N=128 K=p
//                    A = Theta resident in TMEM (TS mode), B = X tile, K-major SWIZZLE_128B
//   R^T  [2S x 128] = score(y, eta)                       epilogue warps, IN PLACE in TMEM:
//                    tcgen05.ld -> registers -> log-density / score -> tcgen05.st
//   G^T  [2S x p]  += R^T[2S x 128] . Xtile[128 x p]      tcgen05.mma kind::tf32, M=128 N=p K=128
//                    A = the eta accumulator columns themselves (TS mode), B = X tile, MN-major
//
// so the residual matrix never touches shared memory, and an epilogue thread owns one particle:
// its per-particle statistics are single registers.
//
// TWO MMA ROWS PER PARTICLE (hi / lo). An M = 128 MMA costs exactly what an M = 64 one does
// (tools/umma_time.cu: 74.4 cycles at N = 128, 42.4 at N = 64, A in TMEM), so the 64 rows an
// S <= 64 problem leaves idle carry the low-order halves of the per-particle operands:
// theta_s = hi + lo with hi = tf32(theta_s), lo = tf32(theta_s - hi), and likewise for the score.
// Row 32q + i (i < 16) of the MMA is the hi half of particle 16q + i, row 32q + 16 + i its lo
// half, so both halves live in the TMEM lane quadrant of epilogue warp q and come back with the
// same register mapping from two tcgen05.ld.16x256b (lane offsets 0 and 16,
// tools/tmem_shape_probe.cu). eta = hi row + lo row carries theta to 22 bits: the systematic
// per-particle error of a TF32-rounded Theta (it does not average out over rows) is gone, at no
// tensor-pipe cost.
//
// Facts measured on B200 that shape this kernel (tools/umma_probe.cu, umma_time.cu,
// umma_ts_probe.cu; numbers in DESIGN.md):
//  * kind::tf32 accepts MN-major operands ONLY in the SWIZZLE_128B_BASE32B layout and K-major
//    operands only in the 16-byte-atom layouts, so the X tile is kept in shared memory as two
//    images, each written by its own TMA load (the second one hits L2: HBM is read once).
//  * one tf32 MMA costs about N/2 + 11 cycles, plus M/4 when A is fetched from shared memory;
//    A-in-TMEM removes that term, and MN-major costs the same as K-major.
//  * MMAs must be issued by an `elect.sync`-elected lane under warp-uniform control flow; a
//    `threadIdx.x == 0` guard makes the compiler wrap every UTCHMMA in a vote loop (~90 cycles).
//  * the tensor core accumulates with truncation: ~1e4 chained accumulations into one TMEM tile
//    bias the sum by ~3e-4 relative. The gradient accumulator is therefore ping-ponged between two
//    TMEM tiles and drained into fp32 registers every kFlush tiles.
//
// Precision mode (stated): X is rounded to nearest-even to TF32 (10-bit mantissa)
// by the TMA unit on its way from HBM to shared memory - an unbiased per-element error that
// averages out over rows; Theta and the scores R enter as hi + lo TF32 pairs (21 mantissa bits,
// see above); products are exact, accumulation is fp32 in TMEM; log-densities, residual
// statistics and reductions are fp32 / fp64 SIMT.
//
//  * register-staged loads cannot stream X fast enough: a warp sustains only ~4-5 outstanding
//    LDG.128, so HBM rate is set by the NUMBER OF WARPS issuing loads (tools/ldg_probe.cu: 8
//    warps/SM 4.0 TB/s, 16 -> 6.6, 32 -> 7.3 TB/s). TMA has no such limit and needs no registers.
//  * TMA with tensor-map data type TFLOAT32 rounds fp32 to tf32 to-nearest-EVEN on its way into
//    shared memory, and SWIZZLE_128B / SWIZZLE_128B_ATOM_32B produce exactly the two operand
//    images, zero-filling rows past the end (tools/tma_probe.cu: 8192/8192 elements incl. ties).
//
// Data movement: one elected lane per ring issues cp.async.bulk.tensor loads. The MN-major ring is
// kMnStages deep and is what keeps HBM requests in flight; the K-major image of a tile is only
// needed until its eta product has retired, so that ring is kKStages deep and its loads hit L2
// (the same rows were fetched moments earlier for the MN-major image). X is read from HBM once.
//
// Warp roles (384 threads): warps 0-3 and 8-11 epilogue (TMEM lane quadrant == warp id % 4; lanes
// 0-15 of each quadrant carry the hi rows of 16 particles, lanes 16-31 their lo rows; warps 0-3
// take tile rows 0-63, warps 8-11 rows 64-127 - the epilogue is the longest per-tile chain, and
// two warps per scheduler also hide the tcgen05.ld latency), warp 4 MN-ring TMA producer, warp 5
// K-ring TMA producer, warp 6 stages (y, live) pairs of each tile, warp 7 allocates TMEM and issues
// every MMA.
//
// Replaces: aten::mv / addmv_ and MvBackward of `X @ theta` (tests/test_mininf.py:11,
// examples/minibatch.md:33) plus the element-wise Normal / Bernoulli / Poisson log_prob chains
// and their autograd twins (mininf/core.py:241), for all S particles in one pass.
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "dense_simt.cuh"

namespace mnf {
namespace tc {

constexpr int kP = 64;          // features handled by this instantiation
constexpr int kNS = 64;         // particle slots; S <= kNS, spare slots carry theta = 0
constexpr int kMmaM = 2 * kNS;  // MMA M: a hi and a lo row per particle slot
constexpr int kTileM = 128;     // rows per tile (MMA N of the eta product, K of the gradient product)
constexpr int kKStages = 2;     // K-major image ring (live until the eta product retires)
constexpr int kMnStages = 4;    // MN-major image ring (live until the gradient product retires)
constexpr int kFlush = 8;       // tiles accumulated in TMEM before the gradient tile is drained
// Developer A/B switches (tools/r2_variants.sh): epilogue warps per TMEM lane quadrant, and whether
// the scores go back as hi / lo pairs (M = 128 gradient product) or as one TF32-rounded row
// (M = 64 gradient product: half the tcgen05.st traffic, half the gradient accumulator traffic).
#ifndef MNF_TC_EPI_WARPS
#define MNF_TC_EPI_WARPS 8
#endif
#ifndef MNF_TC_R_LO
#define MNF_TC_R_LO 0
#endif
constexpr int kEpiWarps = MNF_TC_EPI_WARPS;   // 4: warps 0-3 | 8: warps 0-3 (tile rows 0-63) and 8-11 (rows 64-127)
constexpr int kHalves = kEpiWarps / 4;
constexpr bool kRLo = MNF_TC_R_LO != 0;
static_assert(kEpiWarps == 4 || kEpiWarps == 8, "one or two epilogue warps per TMEM lane quadrant");
constexpr int kWarpMnTma = 4, kWarpKTma = 5, kWarpY = 6, kMmaWarp = 7;
constexpr int kThreads = (kEpiWarps == 4 ? 8 : 12) * 32;

constexpr uint32_t kAtomBytes = kTileM * 128;                 // 128 rows x 128 B (32 fp32)
constexpr uint32_t kXImageBytes = (kP / 32) * kAtomBytes;     // 32 KB per image
constexpr uint32_t kYBytes = 2 * kTileM * 4;                  // interleaved (y, live) pairs per row

constexpr uint32_t kOffK = 0;
constexpr uint32_t kOffMN = kOffK + kKStages * kXImageBytes;
constexpr uint32_t kOffY = kOffMN + kMnStages * kXImageBytes;
constexpr uint32_t kOffBar = kOffY + kMnStages * kYBytes;
constexpr uint32_t kNumBars = 2 * kKStages + 2 * kMnStages + 8;  // + eta_full, r_ready, g_full, g_empty (x2 each)
constexpr uint32_t kOffMisc = kOffBar + 8 * kNumBars;         // tmem slot, counters, per-particle params
constexpr uint32_t kOffGrad = kOffMisc + 64 + kNS * 16;         // drained gradient [kNS][kP + 1] fp32
constexpr uint32_t kOffStat = kOffGrad + kNS * (kP + 1) * 4;    // per-particle statistic exchange [2 halves][2][kNS]
constexpr uint32_t kSmemBytes = kOffStat + 4 * kNS * 4 + 1024 /* alignment slack */;
static_assert(kOffMisc % 16 == 0, "misc block alignment");
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

// tensor memory map (512 columns allocated): two eta^T / R^T tiles, two gradient tiles, Theta
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kColEta = 0;               // + b * kTileM
constexpr uint32_t kColG = 2 * kTileM;        // + gb * kP
constexpr uint32_t kColTheta = kColG + 2 * kP;

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
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
  return pred;
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
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// 2-D tiled TMA load: box (32 features x 128 rows) at (feature c0, row c1) -> shared memory
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem descriptor given as two 32-bit words]
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %4, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns of TMEM <-> 16 registers per thread
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
      "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
      "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}
// 16 lanes x 32 columns spread over all 32 threads (measured mapping, tools/tmem_shape_probe.cu):
// register 4g+j of thread t holds lane t/4 + 8*(j>>1), column 8g + 2*(t%4) + (j&1)
__device__ __forceinline__ void tc_ld_16x256b_x4(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_st_16x256b_x4(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- descriptors (cute/arch/mma_sm100_desc.hpp, cute/atom/mma_traits_sm100.hpp) -------------
// Shared-memory matrix descriptor, version 1 (Blackwell):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4
//   [46,48) version = 1 | [61,64) layout type (2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// Instruction descriptor, kind::tf32, fp32 accumulate:
//   [4,6) D format 1=F32 | [7,10) A format 2=TF32 | [10,13) B format 2=TF32 | [15] A MN-major
//   [16] B MN-major | [17,23) N>>3 | [24,29) M>>4
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// round-to-nearest (ties away) to TF32: the tensor core then only drops zero bits
__device__ __forceinline__ uint32_t rn_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
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
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
  return r;
}

#ifdef MNF_TC_DEBUG
// Debug builds only: per-role cycle counters kept in registers and flushed once per thread.
__device__ float* g_tc_debug = nullptr;   // [16384 floats scratch][int64 cycle counters]
#define TC_DECL() long long tc_cnt__[4] = {0, 0, 0, 0}; long long t0__ = clock64(); const long long tstart__ = t0__
#define TC_T0() t0__ = clock64()
#define TC_ACC(i) do { const long long n__ = clock64(); tc_cnt__[i] += n__ - t0__; t0__ = n__; } while (0)
#define TC_FLUSH(slot0, n, cond) do { if (blockIdx.x == 0 && g_tc_debug && (cond)) { \
    for (int i__ = 0; i__ < (n); ++i__) ((long long*)(g_tc_debug + 16384))[(slot0) + i__] = tc_cnt__[i__]; \
    ((long long*)(g_tc_debug + 16384))[15] = clock64() - tstart__; } } while (0)
#else
#define TC_DECL() do {} while (0)
#define TC_T0() do {} while (0)
#define TC_ACC(i) do {} while (0)
#define TC_FLUSH(slot0, n, cond) do {} while (0)
#endif

struct TileCounters {   // particle-independent sums over the live rows this CTA handled
  float n_live;
  float pad;
  double lgamma_sum;    // Poisson: sum lgamma(y + 1)
};

// partial layout per CTA: [S][ncol], ncol = 1 + kP + 2 (same as the SIMT variant)
template <int FAMILY, bool ICPT>
__global__ void __launch_bounds__(kThreads, 1)
dense_tc_kernel(const __grid_constant__ CUtensorMap map_k, const __grid_constant__ CUtensorMap map_mn,
                mnf_dense_site_t site, const float* __restrict__ z, int S, int D,
                float* __restrict__ partial, uint32_t* __restrict__ status) {
  extern __shared__ uint8_t smem_raw[];
  // swizzled operand images need 1024-byte alignment
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sK = base + kOffK, sMN = base + kOffMN, sY = base + kOffY;
  const uint32_t bars = base + kOffBar;
  const uint32_t bKFull = bars, bKEmpty = bKFull + 8 * kKStages;
  const uint32_t bMnFull = bKEmpty + 8 * kKStages, bMnEmpty = bMnFull + 8 * kMnStages;
  const uint32_t bEtaFull = bMnEmpty + 8 * kMnStages, bRReady = bEtaFull + 16;
  const uint32_t bGFull = bRReady + 16, bGEmpty = bGFull + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbase + kOffMisc);
  TileCounters* counters = reinterpret_cast<TileCounters*>(gbase + kOffMisc + 16);
  DenseParticle* sPar = reinterpret_cast<DenseParticle*>(gbase + kOffMisc + 64);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  const int64_t n_tiles = (site.n_rows + kTileM - 1) / kTileM;
  // tiles owned by this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int64_t my_tiles = (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;

  // ---- one-time setup ----------------------------------------------------------------------
  if (tid == 0) {
    for (int i = 0; i < kKStages; ++i) {
      mbar_init(bKFull + 8 * i, 1);    // arrive.expect_tx of the K-ring producer
      mbar_init(bKEmpty + 8 * i, 1);   // tcgen05.commit after the eta product
    }
    for (int i = 0; i < kMnStages; ++i) {
      mbar_init(bMnFull + 8 * i, 2);   // arrive.expect_tx of the MN-ring producer + the y warp
      mbar_init(bMnEmpty + 8 * i, 1);  // tcgen05.commit after the gradient product
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
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)), "n"(kTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == kWarpMnTma && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_mn) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_k) : "memory");
  }
  for (int s = tid; s < kNS; s += kThreads) {
    DenseParticle dp;
    dp.icpt = 0.f; dp.scale = 1.f; dp.dscale = 0.f;
    if (s < S) {
      dp = dense_particle(site, z + (int64_t)s * D);
      if (FAMILY == MNF_NORMAL && !(dp.scale > 0.0f)) atomicOr(status, MNF_ST_BAD_PARAM);
    }
    sPar[s] = dp;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    // Theta -> TMEM as the A operand of the eta product (M = 128: MMA row m lives on lane m).
    // Lane 32*warp + i (i < 16) holds hi = tf32(theta) of particle 16*warp + i, lane 32*warp + 16 + i
    // its lo = tf32(theta - hi); feature j in column kColTheta + j. Spare particle slots are zero.
    const int s = warp * 16 + (lane & 15);
    const bool owner = s < S;
    const bool lo_row = lane >= 16;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
#pragma unroll
    for (int ch = 0; ch < kP / 16; ++ch) {
      uint32_t v[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float th = owner ? z[(int64_t)s * D + site.theta_lat + ch * 16 + c] : 0.f;
        const uint32_t hi = rn_tf32(th);
        v[c] = lo_row ? rn_tf32(th - __uint_as_float(hi)) : hi;
      }
      tc_st16(tmem + lane_base + kColTheta + ch * 16, v);
    }
    tc_wait_st();
    tc_fence_before();
  }
  __syncthreads();
  tc_fence_after();

  if (warp == kWarpMnTma) {
    // ================= MN-major ring: TMA producer (one elected lane) =========================
    if (elect_one()) {
      for (int64_t k = 0; k < my_tiles; ++k) {
        const int st = (int)(k % kMnStages);
        const int row0 = (int)((blockIdx.x + k * gridDim.x) * kTileM);
        mbar_wait(bMnEmpty + 8 * st, (uint32_t)(((k / kMnStages) & 1) ^ 1));
        mbar_arrive_expect_tx(bMnFull + 8 * st, kXImageBytes);
#pragma unroll
        for (int a = 0; a < kP / 32; ++a)
          tma_load_2d(sMN + (uint32_t)st * kXImageBytes + a * kAtomBytes, &map_mn, a * 32, row0, bMnFull + 8 * st);
      }
    }
    __syncwarp();
  } else if (warp == kWarpKTma) {
    // ================= K-major ring: TMA producer (one elected lane) ==========================
    if (elect_one()) {
      for (int64_t k = 0; k < my_tiles; ++k) {
        const int st = (int)(k % kKStages);
        const int row0 = (int)((blockIdx.x + k * gridDim.x) * kTileM);
        mbar_wait(bKEmpty + 8 * st, (uint32_t)(((k / kKStages) & 1) ^ 1));
        mbar_arrive_expect_tx(bKFull + 8 * st, kXImageBytes);
#pragma unroll
        for (int a = 0; a < kP / 32; ++a)
          tma_load_2d(sK + (uint32_t)st * kXImageBytes + a * kAtomBytes, &map_k, a * 32, row0, bKFull + 8 * st);
      }
    }
    __syncwarp();
  } else if (warp == kWarpY) {
    // ================= y warp: (y, live) pairs of each tile, four rows per lane ===============
    // Values of the next kMnStages tiles are kept in registers so their latency is off the path.
    const bool y_vec = (reinterpret_cast<uintptr_t>(site.y) % 16 == 0) &&
                       (site.mask == nullptr || reinterpret_cast<uintptr_t>(site.mask) % 4 == 0);
    float4 yq[kMnStages];
    uint32_t mq[kMnStages];
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
    for (int i = 0; i < kMnStages; ++i) {
      yq[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      mq[i] = 0;
      if (i < my_tiles) fetch(i, yq[i], mq[i]);
    }
    bool bad_value = false;
    int live_total = 0;          // per-lane sums, combined once after the last tile: fp64 adds and
    double lgam_total = 0.0;     // shuffles are slow enough to matter inside the per-tile loop
    for (int64_t k0 = 0; k0 < my_tiles; k0 += kMnStages) {
#pragma unroll
      for (int i = 0; i < kMnStages; ++i) {
        const int64_t k = k0 + i;
        if (k < my_tiles) {
          mbar_wait(bMnEmpty + 8 * i, (uint32_t)(((k / kMnStages) & 1) ^ 1));
          const float yr[4] = {yq[i].x, yq[i].y, yq[i].z, yq[i].w};
          float yv[4], lv[4];
          float lgam = 0.f;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const bool live = ((mq[i] >> (8 * q)) & 0xFFu) != 0;
            yv[q] = live ? yr[q] : 0.f;
            lv[q] = live ? 1.f : 0.f;
            if (live) {
              ++live_total;
              if (!in_support(FAMILY, yr[q])) bad_value = true;
              if (FAMILY == MNF_POISSON) lgam += log_factorial(yr[q]);
            }
          }
          if (FAMILY == MNF_POISSON) lgam_total += (double)lgam;
          const uint32_t dst = sY + (uint32_t)i * kYBytes + lane * 32;
          sts128(dst, __float_as_uint(yv[0]), __float_as_uint(lv[0]), __float_as_uint(yv[1]), __float_as_uint(lv[1]));
          sts128(dst + 16, __float_as_uint(yv[2]), __float_as_uint(lv[2]), __float_as_uint(yv[3]), __float_as_uint(lv[3]));
          __syncwarp();
          if (lane == 0) mbar_arrive(bMnFull + 8 * i);
          if (k + kMnStages < my_tiles) fetch(k + kMnStages, yq[i], mq[i]);
        }
      }
    }
    if (bad_value) atomicOr(status, MNF_ST_BAD_VALUE);
    // particle-independent sums of this CTA; the epilogue warps read them behind the same barrier
    const float live_sum = warp_sum((float)live_total);
    if (FAMILY == MNF_POISSON) lgam_total = warp_sum(lgam_total);
    if (lane == 0) {
      counters->n_live = live_sum;
      counters->lgamma_sum = lgam_total;
    }
    asm volatile("bar.sync 1, %0;" ::"n"((kEpiWarps + 1) * 32) : "memory");
  } else if (warp == kMmaWarp) {
    // ================= MMA issuer: warp-uniform loop, one elected lane issues ================
    constexpr uint32_t idesc_eta = idesc_tf32(kMmaM, kTileM, 0, 0);   // M=128 N=128, B K-major
    // gradient product: M = 128 over hi and lo score rows, or M = 64 over the hi rows alone (the
    // M = 64 A operand and accumulator live on lanes 0-15 of every quadrant: exactly the hi rows)
    constexpr uint32_t idesc_g = idesc_tf32(kRLo ? kMmaM : kNS, kP, 0, 1);   // N=64, B MN-major
    // descriptor words that do not depend on the stage
    const uint64_t dK = smem_desc(sK, 16, 1024, 2);                 // K-major SWIZZLE_128B
    const uint64_t dMN = smem_desc(sMN, kAtomBytes, 512, 1);        // MN-major SWIZZLE_128B_BASE32B
    const uint32_t dK_lo = (uint32_t)dK, dK_hi = (uint32_t)(dK >> 32);
    const uint32_t dMN_lo = (uint32_t)dMN, dMN_hi = (uint32_t)(dMN >> 32);
    TC_DECL();
    for (int64_t k = 0; k <= my_tiles; ++k) {
      TC_T0();
      if (k < my_tiles) {
        const int kst = (int)(k % kKStages), mst = (int)(k % kMnStages);
        const uint32_t b = (uint32_t)(k & 1);
        mbar_wait(bKFull + 8 * kst, (uint32_t)((k / kKStages) & 1));
        // the MN-major stage carries the (y, live) pairs the epilogue of this tile will read
        mbar_wait(bMnFull + 8 * mst, (uint32_t)((k / kMnStages) & 1));
        TC_ACC(0);   // mma: wait operands
        tc_fence_after();
        if (elect_one()) {
          const uint32_t lo = dK_lo + (uint32_t)kst * (kXImageBytes >> 4);
          const uint32_t d = tmem + kColEta + b * kTileM;
#pragma unroll
          for (int a = 0; a < kP / 32; ++a) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              tc_mma_ts(d, tmem + kColTheta + (a * 4 + ks) * 8,
                        lo + ((a * kAtomBytes + ks * 32) >> 4), dK_hi, idesc_eta, (a | ks) != 0 ? 1u : 0u);
            }
          }
          tc_commit(bEtaFull + 8 * b);
          tc_commit(bKEmpty + 8 * kst);
        }
        __syncwarp();
        TC_ACC(1);   // mma: issue eta
      }
      if (k >= 1) {
        const int64_t kk = k - 1;
        const int mst = (int)(kk % kMnStages);
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
          const uint32_t lo = dMN_lo + (uint32_t)mst * (kXImageBytes >> 4);
          const uint32_t d = tmem + kColG + gb * kP;
          const uint32_t a0 = tmem + kColEta + b * kTileM;
#pragma unroll
          for (int ks = 0; ks < kTileM / 8; ++ks) {
            tc_mma_ts(d, a0 + ks * 8, lo + ks * 64, dMN_hi, idesc_g, (!first || ks > 0) ? 1u : 0u);
          }
          tc_commit(bMnEmpty + 8 * mst);
          if (last) tc_commit(bGFull + 8 * gb);
        }
        __syncwarp();
        TC_ACC(3);   // mma: issue G
      }
    }
    TC_FLUSH(4, 4, lane == 0);
  } else {
    // ================= epilogue warps: quadrant q = warp % 4, column half = warp / 8 ============
    // thread (q, lane < 16) is the owner of particle 16*q + lane for its half of the tile rows
    const int q = warp & 3, half = warp >> 3;
    const int s = q * 16 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    // Statistics: with the 16x256b access shape thread t works on particles 16*q + t/4 (A) and
    // + 8 (B), columns 8g + 2*(t%4) + {0,1} of every 8-column group.
    double stA_total = 0.0, stB_total = 0.0;  // Normal: sum r^2 | others: sum log-density (w/o lgamma)
    double scA_total = 0.0, scB_total = 0.0;  // sum of scores: the intercept gradient (ICPT only)
    const float icptA = ICPT ? sPar[q * 16 + (lane >> 2)].icpt : 0.f;
    const float icptB = ICPT ? sPar[q * 16 + (lane >> 2) + 8].icpt : 0.f;
    // drained gradient tiles live in shared memory, one padded row per particle (conflict-free);
    // this warp owns features [32*half, 32*half + 32) of its 16 particles
    constexpr int kDrainCols = kP / kHalves;    // features drained by this warp
    float* grad_row = reinterpret_cast<float*>(gbase + kOffGrad) + (size_t)(lane < 16 ? s : 0) * (kP + 1) + kDrainCols * half;
    if (lane < 16)
      for (int j = 0; j < kDrainCols; ++j) grad_row[j] = 0.f;
    int64_t n_drained = 0;

    auto drain = [&]() {
      const int64_t grp = n_drained;
      const uint32_t gb = (uint32_t)(grp & 1);
      mbar_wait(bGFull + 8 * gb, (uint32_t)((grp >> 1) & 1));
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < kDrainCols / 32; ++ch) {
        // 32x32b: thread == TMEM lane; lanes 0-15 hold the rows fed by the hi scores of particle
        // 16*q + lane, lanes 16-31 (kRLo only) the rows fed by its lo scores
        uint32_t v[32];
        tc_ld32(tmem + lane_base + kColG + gb * kP + half * kDrainCols + ch * 32, v);
        tc_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          float g = __uint_as_float(v[c]);
          if (kRLo) g += __shfl_down_sync(0xffffffffu, g, 16);
          if (lane < 16) grad_row[ch * 32 + c] += g;
        }
      }
      tc_fence_before();
      mbar_arrive(bGEmpty + 8 * gb);
      ++n_drained;
    };

    // one (row, particle) point: eta = hi row + lo row; the score goes back as a hi / lo pair for
    // the gradient product, and into the running statistic
    auto point = [&](uint32_t& cell, uint32_t& cell_lo, float y, float live, float icpt, float& stat, float& ssum) {
      float eta = __uint_as_float(cell) + __uint_as_float(cell_lo);
      if (ICPT) eta += icpt;
      float score;
      if (FAMILY == MNF_NORMAL) {
        score = fmaf(-live, eta, y);             // live * (y - eta); 1/sigma^2 applied at the end
        stat = fmaf(score, score, stat);
      } else if (FAMILY == MNF_BERNOULLI_LOGITS) {
        const float e = __expf(-fabsf(eta));
        const float inv = __fdividef(1.0f, 1.0f + e);
        const float sig = eta >= 0.f ? inv : e * inv;
        score = live * (y - sig);
        stat += live * (y * eta - (fmaxf(eta, 0.f) + __logf(1.0f + e)));   // softplus, abs. error ~1e-7
      } else {
        const float rate = __expf(eta);
        score = live * (y - rate);
        stat += live * fmaf(y, eta, -rate);
      }
      if (ICPT) ssum += score;
      const uint32_t hi = rn_tf32(score);
      cell = hi;
      if (kRLo) cell_lo = rn_tf32(score - __uint_as_float(hi));
    };
    // 32 columns (tile rows): thread t touches rows 32ch + 8g + 2(t%4) + {0,1}; their (y, live)
    // pairs are one 16-byte shared-memory word
    auto process = [&](uint32_t (&v)[16], uint32_t (&l)[16], const float4* yl, int ch, float& sa, float& sb,
                       float& ra, float& rb) {
      float4 w[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) w[g] = yl[16 * ch + 4 * g + (lane & 3)];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        point(v[4 * g + 0], l[4 * g + 0], w[g].x, w[g].y, icptA, sa, ra);
        point(v[4 * g + 1], l[4 * g + 1], w[g].z, w[g].w, icptA, sa, ra);
        point(v[4 * g + 2], l[4 * g + 2], w[g].x, w[g].y, icptB, sb, rb);
        point(v[4 * g + 3], l[4 * g + 3], w[g].z, w[g].w, icptB, sb, rb);
      }
    };

    constexpr uint32_t kLoRows = 16u << 16;   // TMEM lane offset of the lo rows inside a quadrant
    TC_DECL();
    for (int64_t k = 0; k < my_tiles; ++k) {
      TC_T0();
      const uint32_t b = (uint32_t)(k & 1);
      const float4* yl = reinterpret_cast<const float4*>(gbase + kOffY + (size_t)(k % kMnStages) * kYBytes);
      // the gradient group that ended two tiles ago has been issued (its R tile was handed over
      // two iterations back), so waiting for its commit cannot deadlock
      if (k >= 2 && ((k - 2) % kFlush) == kFlush - 1) drain();
      mbar_wait(bEtaFull + 8 * b, (uint32_t)((k >> 1) & 1));
      TC_ACC(0);   // epi: wait eta_full (+ drain)
      tc_fence_after();
      constexpr int kChunks = 4 / kHalves;      // 32-column chunks of the tile handled by this warp
      const uint32_t t_eta = tmem + lane_base + kColEta + b * kTileM + 32 * kChunks * half;
      // software pipeline over the chunks: the next chunk's tcgen05.ld pair (hi rows, lo rows) is in
      // flight while the current one is processed; R^T replaces eta^T in place (A operand of the G product)
      float sa = 0.f, sb = 0.f, ra = 0.f, rb = 0.f;
      uint32_t v[2][16], l[2][16];
      tc_ld_16x256b_x4(t_eta, v[0]);
      tc_ld_16x256b_x4(t_eta + kLoRows, l[0]);
#pragma unroll
      for (int ch = 0; ch < kChunks; ++ch) {
        tc_wait_ld();
        if (ch + 1 < kChunks) {
          tc_ld_16x256b_x4(t_eta + 32 * (ch + 1), v[(ch + 1) & 1]);
          tc_ld_16x256b_x4(t_eta + kLoRows + 32 * (ch + 1), l[(ch + 1) & 1]);
        }
        process(v[ch & 1], l[ch & 1], yl, kChunks * half + ch, sa, sb, ra, rb);
        tc_st_16x256b_x4(t_eta + 32 * ch, v[ch & 1]);
        if (kRLo) tc_st_16x256b_x4(t_eta + kLoRows + 32 * ch, l[ch & 1]);
      }
      tc_wait_st();
      tc_fence_before();
      mbar_arrive(bRReady + 8 * b);
      stA_total += (double)sa;
      stB_total += (double)sb;
      if (ICPT) {
        scA_total += (double)ra;
        scB_total += (double)rb;
      }
      TC_ACC(1);   // epi: compute
    }
    TC_FLUSH(8, 2, tid == 0);
    // drain the gradient groups still in tensor memory (at most the last two)
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
    float* s_stat = reinterpret_cast<float*>(gbase + kOffStat) + half * 2 * kNS;   // [half][2][kNS]
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
    const bool owner_thread = half == 0 && lane < 16;
    float st0 = 0.f, sc0 = 0.f;
    if (owner_thread) {
#pragma unroll
      for (int h = 0; h < kHalves; ++h) {
        st0 += s_all[2 * kNS * h + s];
        if (ICPT) sc0 += s_all[2 * kNS * h + kNS + s];
      }
    }
    grad_row -= kDrainCols * half;

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

}  // namespace tc
}  // namespace mnf
