// Element-wise site sweep: sites whose value is a long observed vector and whose parameters are
// scalar links  T(A_s + B_s * x_i)  of scalar latents (the missing-observations config:
// counts ~ Poisson(exp(a + b x)), w ~ Normal(c + d x, sigma), both masked;
// examples/missing-observations.md:131, mininf/core.py:231-239,262-265).
//
// Mapping: particles live on LANES (lane l owns particles l, l+32, ...), elements are loaded
// coalesced 32 at a time (one per lane) together with their data-only terms (lgamma(v+1)), and
// then broadcast one by one with warp shuffles. Every lane therefore accumulates the sums of ITS
// particles privately - no atomics, no cross-lane reduction in the hot loop - and masked-out
// elements are skipped warp-uniformly (no compaction copy, unlike `data[mask].sum()`).
// Up to kMaxSites sites of equal length are fused so shared covariates are read once.
#pragma once

#include <type_traits>

#include "common.cuh"

namespace mnf {

constexpr int kSweepThreads = 256;
constexpr int kSweepWarps = kSweepThreads / 32;

template <int NSITES>
struct SweepArgs {
  mnf_site_t site[NSITES];
};

// log-density and d/du0, d/du1 for one element/particle; c = data-only term (Poisson lgamma)
__device__ __forceinline__ void sweep_point(int family, float v, float c, float u0, int t0, float u1,
                                            int t1, float& lp, float& du0, float& du1, bool& bad) {
  du1 = 0.0f;
  if (family == MNF_POISSON) {
    if (t0 == MNF_T_EXP) {  // rate = exp(u): log rate = u exactly
      const float rate = expf(u0);
      lp = fmaf(v, u0, -rate) - c;
      du0 = v - rate;
    } else {
      lp = xlogy(v, u0) - u0 - c;
      du0 = v == 0.0f ? -1.0f : v / u0 - 1.0f;
      bad |= !(u0 >= 0.0f);
    }
  } else if (family == MNF_NORMAL) {
    const float sigma = t1 == MNF_T_EXP ? expf(u1) : u1;
    const float loc = t0 == MNF_T_EXP ? expf(u0) : u0;
    const float inv = 1.0f / sigma;
    const float r = (v - loc) * inv;
    lp = -0.5f * r * r - logf(sigma) - kLogSqrt2Pi;
    du0 = r * inv * (t0 == MNF_T_EXP ? loc : 1.0f);
    du1 = (r * r - 1.0f) * inv * (t1 == MNF_T_EXP ? sigma : 1.0f);
    bad |= !(sigma > 0.0f);
  } else {
    const float p0 = t0 == MNF_T_EXP ? expf(u0) : u0;
    const float p1 = t1 == MNF_T_EXP ? expf(u1) : u1;
    const Dens d = density(family, v, p0, p1, true);
    lp = d.lp;
    du0 = d.d0 * (t0 == MNF_T_EXP ? p0 : 1.0f);
    du1 = d.d1 * (t1 == MNF_T_EXP ? p1 : 1.0f);
    bad |= d.bad_param;
  }
}

// Normal site whose scale is a per-particle constant: sigma, 1/sigma, log sigma and d sigma/du
// are hoisted out of the element loop (the common case: `Normal(c + d*x, sigma)`).
struct ScaleConst {
  float inv;     // 1 / sigma
  float logs;    // log sigma + log sqrt(2 pi)
  float dsig;    // d sigma / d u (sigma for the exp link, 1 for identity)
};

__device__ __forceinline__ void normal_point_hoisted(float v, float u0, int t0, const ScaleConst& sc,
                                                     float& lp, float& du0, float& du1) {
  const float loc = t0 == MNF_T_EXP ? expf(u0) : u0;
  const float r = (v - loc) * sc.inv;
  lp = fmaf(-0.5f * r, r, -sc.logs);
  du0 = r * sc.inv * (t0 == MNF_T_EXP ? loc : 1.0f);
  du1 = fmaf(r, r, -1.0f) * sc.inv * sc.dsig;
}

// partial layout per CTA: [S][ncol], ncol = 1 + 4*NSITES: col 0 scaled log-density, then per site
// the scaled sums of (du0, du0*x0, du1, du1*x1)
template <int NSITES, int Q>
__global__ void __launch_bounds__(kSweepThreads, 3)
site_sweep_kernel(SweepArgs<NSITES> args, const float* __restrict__ z, int S, int D,
                  float* __restrict__ partial, uint32_t* __restrict__ status) {
  constexpr int NC = 1 + 4 * NSITES;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n = args.site[0].numel;

  // latent scalars of this lane's particles
  float A[NSITES][2][Q], B[NSITES][2][Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int s = lane + 32 * q;
    const float* zs = z + (int64_t)(s < S ? s : 0) * D;
#pragma unroll
    for (int i = 0; i < NSITES; ++i)
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        const mnf_link_t& L = args.site[i].param[p];
        A[i][p][q] = L.a_const + (L.a_lat >= 0 ? zs[L.a_lat] : 0.0f);
        B[i][p][q] = L.b_const + (L.b_lat >= 0 ? zs[L.b_lat] : 0.0f);
      }
  }
  // fp64 running sums live in shared memory (touched once per 32-element chunk), laid out as the
  // final cross-warp reduction reads them: [warp][particle slot][column]
  extern __shared__ double s_acc[];
  double* my_acc = s_acc + ((size_t)warp * Q * 32 + lane) * NC;     // + q * 32 * NC + c
#pragma unroll
  for (int q = 0; q < Q; ++q)
#pragma unroll
    for (int c = 0; c < NC; ++c) my_acc[(size_t)q * 32 * NC + c] = 0.0;

  bool bad_param = false, bad_value = false;
  // Normal sites with an element-independent scale: hoist its derived constants per particle
  bool hoisted[NSITES];
  ScaleConst sconst[NSITES][Q];
#pragma unroll
  for (int k = 0; k < NSITES; ++k) {
    const mnf_link_t& L = args.site[k].param[1];
    hoisted[k] = args.site[k].family == MNF_NORMAL && L.x == nullptr;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const float u = A[k][1][q] + B[k][1][q];           // x == NULL means x == 1
      const float sigma = L.transform == MNF_T_EXP ? expf(u) : u;
      sconst[k][q].inv = 1.0f / sigma;
      sconst[k][q].logs = logf(sigma) + kLogSqrt2Pi;
      sconst[k][q].dsig = L.transform == MNF_T_EXP ? sigma : 1.0f;
      if (hoisted[k] && !(sigma > 0.0f) && lane + 32 * q < S) bad_param = true;
    }
  }
  const int64_t n_chunks = (n + 31) / 32;
  const int64_t warp_global = (int64_t)blockIdx.x * kSweepWarps + warp;
  const int64_t warps_total = (int64_t)gridDim.x * kSweepWarps;

  // per-warp staging of a chunk's element data: one float4 (value, data-only term, x0, x1) per
  // site and element, read back as 16-byte broadcasts (one LDS.128 instead of four shuffles)
  float4* stage = reinterpret_cast<float4*>(s_acc + (size_t)kSweepWarps * Q * 32 * NC) + (size_t)warp * NSITES * 32;

  for (int64_t chunk = warp_global; chunk < n_chunks; chunk += warps_total) {
    const int64_t i = chunk * 32 + lane;
    const bool inb = i < n;
    float ev[NSITES], ec[NSITES], ex0[NSITES], ex1[NSITES];
    uint32_t live_bits[NSITES];
#pragma unroll
    for (int k = 0; k < NSITES; ++k) {
      const mnf_site_t& st = args.site[k];
      bool live = inb && (st.mask == nullptr || st.mask[i] != 0);
      ev[k] = live ? __ldg(st.value + i) : 0.0f;
      ex0[k] = (inb && st.param[0].x != nullptr) ? __ldg(st.param[0].x + (int64_t)st.param[0].x_stride * i) : 1.0f;
      ex1[k] = (inb && st.param[1].x != nullptr) ? __ldg(st.param[1].x + (int64_t)st.param[1].x_stride * i) : 1.0f;
      ec[k] = 0.0f;
      if (live) {
        if (!in_support(st.family, ev[k])) bad_value = true;
        if (st.family == MNF_POISSON) ec[k] = log_factorial(ev[k]);
      }
      live_bits[k] = __ballot_sync(0xffffffffu, live);
      stage[k * 32 + lane] = make_float4(ev[k], ec[k], ex0[k], ex1[k]);
    }
    __syncwarp();
    float part[NC][Q];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int q = 0; q < Q; ++q) part[c][q] = 0.0f;

#pragma unroll 2
    for (int e = 0; e < 32; ++e) {
#pragma unroll
      for (int k = 0; k < NSITES; ++k) {
        if (!((live_bits[k] >> e) & 1u)) continue;  // warp-uniform
        const mnf_site_t& st = args.site[k];
        const float4 el = stage[k * 32 + e];
        const float v = el.x, c = el.y, x0 = el.z, x1 = el.w;
        const float w = (float)st.scale;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const float u0 = fmaf(B[k][0][q], x0, A[k][0][q]);
          const float u1 = fmaf(B[k][1][q], x1, A[k][1][q]);
          float lp, du0, du1;
          if (hoisted[k])
            normal_point_hoisted(v, u0, st.param[0].transform, sconst[k][q], lp, du0, du1);
          else
            sweep_point(st.family, v, c, u0, st.param[0].transform, u1, st.param[1].transform, lp,
                        du0, du1, bad_param);
          part[0][q] = fmaf(w, lp, part[0][q]);
          du0 *= w;
          du1 *= w;
          part[1 + 4 * k + 0][q] += du0;
          part[1 + 4 * k + 1][q] = fmaf(du0, x0, part[1 + 4 * k + 1][q]);
          part[1 + 4 * k + 2][q] += du1;
          part[1 + 4 * k + 3][q] = fmaf(du1, x1, part[1 + 4 * k + 3][q]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int c = 0; c < NC; ++c) my_acc[(size_t)q * 32 * NC + c] += (double)part[c][q];
    __syncwarp();   // the staging slots are rewritten by the next chunk
  }

  // cross-warp reduction in fixed order, then one partial block per CTA
  __syncthreads();
  float* out = partial + (size_t)blockIdx.x * S * NC;
  for (int idx = threadIdx.x; idx < S * NC; idx += kSweepThreads) {
    const int s = idx / NC, c = idx % NC;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kSweepWarps; ++w) t += s_acc[((size_t)w * Q * 32 + s) * NC + c];
    out[idx] = (float)t;
  }
  // bad parameters observed by lanes that carry no particle do not count
  if (bad_value) atomicOr(status, MNF_ST_BAD_VALUE);
  if (bad_param && lane < S) atomicOr(status, MNF_ST_BAD_PARAM);
}

template <int NSITES, int Q>
inline size_t site_sweep_smem_bytes() {
  return sizeof(double) * kSweepWarps * Q * 32 * (1 + 4 * NSITES) +     // running sums
         sizeof(float4) * kSweepWarps * NSITES * 32;                    // element staging
}

// ---------------------------------------------------------------------------------------------
// Specialised single-site sweeps for the two hot site kinds of the missing-observations config
// (counts ~ Poisson(exp(a + b x)), w ~ Normal(c + d x, sigma), both masked). The generic kernel
// above spends ~40 instructions per (element, particle); here everything that does not depend on
// the particle is summed once per element.
//
//   Poisson(exp(A_s + B_s x)):  log p = A_s V + B_s Vx - R_s - C,  d/dA = V - R_s,  d/dB = Vx - Rx_s
//       with the data-only sums V = sum v, Vx = sum v x, C = sum log v! over the live elements and
//       the irreducible per-particle sums R_s = sum rate, Rx_s = sum rate x  (one ex2 + four FMA-pipe
//       instructions per (element, particle): the kernel is bound by the MUFU pipe).
//   Normal(A_s + B_s x, sigma_s): the residual sums are polynomials in (A_s, B_s) of SIX data-only
//       sufficient statistics n, Sx, Sxx, Sv, Svx, Svv:
//           T1 = sum r   = Sv  - A n  - B Sx          r = v - A - B x
//           Tx = sum r x = Svx - A Sx - B Sxx
//           T2 = sum r^2 = Svv - A Sv - B Svx - A T1 - B Tx
//       log p = -T2 / (2 sigma^2) - n log(sigma sqrt(2 pi)), d/dA = T1/sigma^2, d/dB = Tx/sigma^2,
//       d/dsigma = (T2/sigma^2 - n)/sigma.
//       So the sweep does NO per-particle work: one HBM-bound pass (value 4 + covariate 4 + mask 1
//       bytes per element) whatever S is. Products and sums are carried in fp64 (fp32 x fp32 is
//       exact in fp64), so the cancellation in T2 costs log10(Svv / T2) of 16 digits.
// ---------------------------------------------------------------------------------------------
constexpr int kFastNone = -1, kFastPoissonExp = 0, kFastNormalId = 1;

inline int site_fast_kind(const mnf_site_t& st) {
  if (st.family == MNF_POISSON && st.param[0].transform == MNF_T_EXP) return kFastPoissonExp;
  if (st.family == MNF_NORMAL && st.param[0].transform == MNF_T_ID && st.param[1].x == nullptr)
    return kFastNormalId;
  return kFastNone;
}

__device__ __forceinline__ float ex2_approx(float t) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
  return r;
}

// ---- Poisson(exp(A_s + B_s x)) -------------------------------------------------------------------
// Particles on lanes (lane l owns particles l, l+32, ...). A warp walks chunks of 32 elements (one
// per lane, coalesced), adds the data-only terms, and appends the covariates of the LIVE elements
// to a 64-entry ring in shared memory (ballot + popc compaction). Whenever 32 are queued they are
// consumed by a branch-free, fully unrolled loop (LDS.128 broadcasts, two FMAs + ex2 + add + FMA
// per particle), so masked-out elements cost nothing in the hot loop and it has no control flow.
// exp(u) = 2^(u log2 e): log2 e is folded into the per-particle constants in fp64 and split as
// hi + lo, the lo part of A is applied once at the end as a factor on R_s and Rx_s.
constexpr int pois_min_blocks(int q) { return q <= 2 ? 3 : 2; }   // 80 / 128 registers per thread

template <int Q>
__global__ void __launch_bounds__(kSweepThreads, pois_min_blocks(Q))
poisson_exp_kernel(mnf_site_t st, const float* __restrict__ z, int S, int D, float* __restrict__ partial,
                   uint32_t* __restrict__ status, const uint32_t* __restrict__ need_exact) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n = st.numel;
  const mnf_link_t L0 = st.param[0];
  // The moment path below already added this site to the accumulator: hand the reduction zeros.
  if (need_exact != nullptr && *need_exact == 0u) {
    for (int idx = threadIdx.x; idx < S * 5; idx += kSweepThreads) partial[(size_t)blockIdx.x * S * 5 + idx] = 0.0f;
    return;
  }

  extern __shared__ double s_pois[];
  double* s_sums = s_pois;                                          // [warp][Q*32][2]
  double* s_elem = s_pois + (size_t)kSweepWarps * Q * 32 * 2;       // [warp][4]: V, Vx, C, n
  float* s_ring = reinterpret_cast<float*>(s_elem + kSweepWarps * 4);   // [warp][64], 16-byte aligned
  float* s_logfact = s_ring + kSweepWarps * 64;                     // [64] copy of the constant table
  float* ring = s_ring + warp * 64;
  if (threadIdx.x < 64) s_logfact[threadIdx.x] = kLogFactorial[threadIdx.x];
  __syncthreads();

  float A2[Q], Bh[Q], Bl[Q];
  double corr[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int s = lane + 32 * q;
    const float* zs = z + (int64_t)(s < S ? s : 0) * D;
    const float A = L0.a_const + (L0.a_lat >= 0 ? zs[L0.a_lat] : 0.0f);
    const float B = L0.b_const + (L0.b_lat >= 0 ? zs[L0.b_lat] : 0.0f);
    const double a2 = (double)A * 1.4426950408889634074, b2 = (double)B * 1.4426950408889634074;
    A2[q] = (float)a2;
    corr[q] = exp2(a2 - (double)A2[q]);
    Bh[q] = (float)b2;
    Bl[q] = (float)(b2 - (double)Bh[q]);
  }

  double run[Q][2];
#pragma unroll
  for (int q = 0; q < Q; ++q) { run[q][0] = 0.0; run[q][1] = 0.0; }
  float r0[Q], r1[Q];                          // fp32 running sums, flushed to `run` every kFlush groups
#pragma unroll
  for (int q = 0; q < Q; ++q) { r0[q] = 0.0f; r1[q] = 0.0f; }
  double e_v = 0.0, e_vx = 0.0, e_c = 0.0;     // this lane's share of the per-element sums
  float f_v = 0.0f, f_vx = 0.0f, f_c = 0.0f;   // fp32 staging of the same, flushed every kFlush chunks
  int e_n = 0;
  bool bad_value = false;
  int head = 0, fill = 0;                      // ring state, warp-uniform
  int groups = 0, chunks = 0;
  constexpr int kFlush = 8;                    // 256 rates / 8 elements per fp32 sum: error ~1e-7 * sqrt(256)

  const int64_t n_chunks = (n + 31) / 32;
  const int64_t warp_global = (int64_t)blockIdx.x * kSweepWarps + warp;
  const int64_t warps_total = (int64_t)gridDim.x * kSweepWarps;
  // software pipeline: the loads of the next kAhead chunks are in flight while the ring is
  // consumed (an iteration that finds fewer than 32 queued covariates is much shorter than the
  // HBM latency, so one chunk of lookahead is not enough)
  constexpr int kAhead = 3;
  float p_v[kAhead], p_x[kAhead];
  uint32_t p_m[kAhead];       // raw mask byte: it is only LOOKED at when the chunk is consumed
  auto fetch = [&](int64_t chunk, float& fv, float& fx, uint32_t& fm) {
    const int64_t i = chunk * 32 + lane;
    const bool inb = chunk < n_chunks && i < n;
    fm = inb ? (st.mask == nullptr ? 1u : (uint32_t)__ldg(st.mask + i)) : 0u;
    fv = inb ? __ldg(st.value + i) : 0.0f;
    fx = (inb && L0.x != nullptr) ? __ldg(L0.x + (int64_t)L0.x_stride * i) : 1.0f;
  };
#pragma unroll
  for (int k = 0; k < kAhead; ++k) fetch(warp_global + k * warps_total, p_v[k], p_x[k], p_m[k]);
  for (int64_t chunk = warp_global; chunk < n_chunks; chunk += warps_total) {
    const float v = p_v[0], x = p_x[0];
    const bool live = p_m[0] != 0u;
#pragma unroll
    for (int k = 0; k + 1 < kAhead; ++k) { p_v[k] = p_v[k + 1]; p_x[k] = p_x[k + 1]; p_m[k] = p_m[k + 1]; }
    fetch(chunk + kAhead * warps_total, p_v[kAhead - 1], p_x[kAhead - 1], p_m[kAhead - 1]);
    if (live) {
      ++e_n;
      f_v += v;
      f_vx = fmaf(v, x, f_vx);
      if (v >= 0.0f && v < 64.0f && v == floorf(v)) {
        f_c += s_logfact[(int)v];
      } else {
        if (!in_support(MNF_POISSON, v)) bad_value = true;
        f_c += log_factorial(v);
      }
    }
    if (++chunks == kFlush) {
      e_v += (double)f_v; e_vx += (double)f_vx; e_c += (double)f_c;
      f_v = 0.0f; f_vx = 0.0f; f_c = 0.0f;
      chunks = 0;
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, live);
    if (live) ring[(head + fill + __popc(bits & ((1u << lane) - 1u))) & 63] = x;
    fill += __popc(bits);
    __syncwarp();
    if (fill >= 32) {
      const float4* b4 = reinterpret_cast<const float4*>(ring + head);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 xs = b4[j];
        const float xe[4] = {xs.x, xs.y, xs.z, xs.w};
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            const float rate = ex2_approx(fmaf(Bh[q], xe[t], fmaf(Bl[q], xe[t], A2[q])));
            r0[q] += rate;
            r1[q] = fmaf(rate, xe[t], r1[q]);
          }
      }
      if (++groups == kFlush) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          run[q][0] += (double)r0[q]; run[q][1] += (double)r1[q];
          r0[q] = 0.0f; r1[q] = 0.0f;
        }
        groups = 0;
      }
      head ^= 32;
      fill -= 32;
      __syncwarp();   // consumed slots are rewritten by the next append
    }
  }
  e_v += (double)f_v; e_vx += (double)f_vx; e_c += (double)f_c;
  // what is left in the ring (< 32 covariates)
  for (int e = 0; e < fill; ++e) {
    const float xe = ring[(head + e) & 63];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const float rate = ex2_approx(fmaf(Bh[q], xe, fmaf(Bl[q], xe, A2[q])));
      r0[q] += rate;
      r1[q] = fmaf(rate, xe, r1[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < Q; ++q) { run[q][0] += (double)r0[q]; run[q][1] += (double)r1[q]; }

#pragma unroll
  for (int q = 0; q < Q; ++q) {
    s_sums[((size_t)warp * Q * 32 + q * 32 + lane) * 2 + 0] = run[q][0] * corr[q];
    s_sums[((size_t)warp * Q * 32 + q * 32 + lane) * 2 + 1] = run[q][1] * corr[q];
  }
  e_v = warp_sum(e_v);
  e_vx = warp_sum(e_vx);
  e_c = warp_sum(e_c);
  const double e_cnt = warp_sum((double)e_n);
  if (lane == 0) {
    s_elem[warp * 4 + 0] = e_v;
    s_elem[warp * 4 + 1] = e_vx;
    s_elem[warp * 4 + 2] = e_c;
    s_elem[warp * 4 + 3] = e_cnt;
  }
  __syncthreads();

  // one particle per thread: combine the warps in fixed order and finish the closed form
  for (int s = threadIdx.x; s < S; s += kSweepThreads) {
    double t[2] = {0.0, 0.0}, el[4] = {0.0, 0.0, 0.0, 0.0};
    for (int w = 0; w < kSweepWarps; ++w) {
      t[0] += s_sums[((size_t)w * Q * 32 + s) * 2 + 0];
      t[1] += s_sums[((size_t)w * Q * 32 + s) * 2 + 1];
#pragma unroll
      for (int a = 0; a < 4; ++a) el[a] += s_elem[w * 4 + a];
    }
    const float* zs = z + (int64_t)s * D;
    const double a_s = (double)(L0.a_const + (L0.a_lat >= 0 ? zs[L0.a_lat] : 0.0f));
    const double b_s = (double)(L0.b_const + (L0.b_lat >= 0 ? zs[L0.b_lat] : 0.0f));
    float* out = partial + ((size_t)blockIdx.x * S + s) * 5;
    const double w = st.scale;
    out[0] = (float)(w * (a_s * el[0] + b_s * el[1] - t[0] - el[2]));
    out[1] = (float)(w * (el[0] - t[0]));
    out[2] = (float)(w * (el[1] - t[1]));
    out[3] = 0.0f;
    out[4] = 0.0f;
  }
  if (bad_value) atomicOr(status, MNF_ST_BAD_VALUE);
}

template <int Q>
inline size_t poisson_exp_smem_bytes() {
  return sizeof(double) * kSweepWarps * Q * 32 * 2 + sizeof(double) * kSweepWarps * 4 +
         sizeof(float) * kSweepWarps * 64 + sizeof(float) * 64;
}

// ---- Poisson(exp(A_s + B_s x)) through data-only Chebyshev moments ------------------------------
// The irreducible per-particle sums of the kernel above, R_s = sum_i exp(A_s + B_s x_i) and
// Rx_s = sum_i x_i exp(A_s + B_s x_i), are the moment generating function of the covariate and its
// derivative. With t = (x - mid) / half in [-1, 1] over the live elements,
//       exp(B half t) = I_0(B half) + 2 sum_{j>=1} I_j(B half) T_j(t)          (Jacobi-Anger)
// so   R_s  = exp(A_s + B_s mid) * sum_j w_j(B_s half) mu_j,    mu_j = sum_i T_j(t_i)   (data only)
//      Rx_s = mid R_s + half exp(A_s + B_s mid) * sum_j w_j'(B_s half) mu_j,  I_j' = (I_{j-1} + I_{j+1}) / 2.
// The series converges like (B half / 2)^j / j!: kChebMoments = 33 terms leave a truncation below
// 1e-13 for |B| half <= 8. The sweep therefore does NO per-particle work: a range pass (5 bytes
// per element) and a moment pass (value 4 + covariate 4 + mask 1 bytes, one FFMA + one FADD per
// moment and element) whatever S is - the MUFU bound of the kernel above disappears. The finish
// kernel evaluates the modified Bessel functions per particle in fp64 and CHECKS the expansion
// (truncation bound and the cancellation sum_j |w_j mu_j| / |sum_j w_j mu_j|) from the data; if any
// particle fails the check it raises `need_exact` and the kernel above runs instead (launched
// unconditionally, it returns at once when the flag is clear), so the result never depends on
// the expansion being applicable.
constexpr int kChebMoments = 33;                   // T_0 .. T_32
constexpr int kChebCols = kChebMoments + 4;        // + V, Vx, C, n
constexpr int kChebThreads = 256;
constexpr double kChebMaxArg = 12.0;               // |B| half beyond which the exact kernel runs
constexpr double kChebMaxCancel = 1024.0;          // moments carry ~1e-9 relative error
constexpr double kChebMaxTrunc = 1e-9;

struct ChebRange {
  float mid, inv;     // t = (x - mid) * inv
  float lo, hi;       // the covariate range it was built from
  bool cached;        // taken from the range slot (not from this step's range pass)
};

// The covariate range of a site, kept across steps in library-owned device memory (site.cu keys the
// slots by the covariate / mask addresses and the element count). The covariate is data: it does not
// change from step to step, so re-reading all of it every step just to find its range was 5 of the
// 23 bytes per element this path moved. A valid slot makes the range pass return at once; the moment
// pass then verifies the slot against every live element it reads (max |t|), and the finish kernel
// drops the slot - and hands the step to the exact kernel - if the data no longer fit (a plan
// re-bound to other rows), or refreshes it if the data use less than half of it.
struct RangeSlot {
  float lo, hi;
  uint32_t valid;
  uint32_t pad;
};

// every CTA reduces the range partials the same way => identical (mid, inv) everywhere
__device__ __forceinline__ ChebRange cheb_range(const float* __restrict__ range_partial, int n_range, float* s_tmp,
                                                const RangeSlot* __restrict__ slot = nullptr) {
  float lo = INFINITY, hi = -INFINITY;
  const bool cached = slot != nullptr && *reinterpret_cast<const volatile uint32_t*>(&slot->valid) != 0u;
  if (cached) {
    lo = *reinterpret_cast<const volatile float*>(&slot->lo);
    hi = *reinterpret_cast<const volatile float*>(&slot->hi);
    n_range = 0;
  }
  for (int b = threadIdx.x; b < n_range; b += blockDim.x) {
    lo = fminf(lo, range_partial[2 * b]);
    hi = fmaxf(hi, range_partial[2 * b + 1]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) { s_tmp[2 * warp] = lo; s_tmp[2 * warp + 1] = hi; }
  __syncthreads();
  lo = INFINITY; hi = -INFINITY;
  for (int w = 0; w < nw; ++w) { lo = fminf(lo, s_tmp[2 * w]); hi = fmaxf(hi, s_tmp[2 * w + 1]); }
  __syncthreads();
  ChebRange r;
  r.lo = lo; r.hi = hi; r.cached = cached;
  if (!(lo <= hi)) { r.mid = 0.0f; r.inv = 0.0f; return r; }      // no live element
  r.mid = 0.5f * lo + 0.5f * hi;
  const float half = fmaxf(hi - r.mid, r.mid - lo);
  r.inv = half > 0.0f ? 1.0f / half : 0.0f;
  return r;
}

// min / max of the covariate over the live elements; 128-bit loads (the launcher checks alignment)
__global__ void __launch_bounds__(kChebThreads)
poisson_range_kernel(mnf_site_t st, float* __restrict__ range_partial, const RangeSlot* __restrict__ slot) {
  if (slot != nullptr && *reinterpret_cast<const volatile uint32_t*>(&slot->valid) != 0u) return;   // range known
  const int64_t n = st.numel;
  const float* __restrict__ xp = st.param[0].x;
  const uint8_t* __restrict__ mask = st.mask;
  const int64_t tid = (int64_t)blockIdx.x * kChebThreads + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * kChebThreads;
  float lo = INFINITY, hi = -INFINITY;
  auto take = [&](float x, bool live) {
    if (live) { lo = fminf(lo, x); hi = fmaxf(hi, x); }
  };
  const int64_t n4 = n >> 2;
  for (int64_t g = tid; g < n4; g += nth) {
    const float4 x4 = __ldg(reinterpret_cast<const float4*>(xp) + g);
    const uint32_t m4 = mask != nullptr ? __ldg(reinterpret_cast<const uint32_t*>(mask) + g) : 0x01010101u;
    take(x4.x, (m4 & 0x000000ffu) != 0);
    take(x4.y, (m4 & 0x0000ff00u) != 0);
    take(x4.z, (m4 & 0x00ff0000u) != 0);
    take(x4.w, (m4 & 0xff000000u) != 0);
  }
  const int64_t i = (n4 << 2) + tid;
  if (i < n) take(__ldg(xp + i), mask == nullptr || __ldg(mask + i) != 0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __shared__ float s_r[2 * kChebThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_r[2 * warp] = lo; s_r[2 * warp + 1] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kChebThreads / 32; ++w) { lo = fminf(lo, s_r[2 * w]); hi = fmaxf(hi, s_r[2 * w + 1]); }
    range_partial[2 * blockIdx.x] = lo;
    range_partial[2 * blockIdx.x + 1] = hi;
  }
}

// Number of moments the expansion needs for the largest |B_s| half of this call: the smallest of
// 13 / 17 / 21 / 25 / 33 with 4 I_{J+1}(arg) exp(arg) <= 1e-10 (truncation relative to the smallest
// possible rate sum), evaluated the same way by the moment and the finish kernel.
__device__ __forceinline__ int cheb_moments_needed(const mnf_link_t& L0, const float* __restrict__ z, int S, int D,
                                                   float inv, float* s_tmp) {
  float bmax = 0.0f;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const float b = L0.b_const + (L0.b_lat >= 0 ? z[(int64_t)s * D + L0.b_lat] : 0.0f);
    bmax = b == b ? fmaxf(bmax, fabsf(b)) : INFINITY;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bmax = fmaxf(bmax, __shfl_xor_sync(0xffffffffu, bmax, o));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) s_tmp[warp] = bmax;
  __syncthreads();
  bmax = 0.0f;
  for (int w = 0; w < nw; ++w) bmax = fmaxf(bmax, s_tmp[w]);
  __syncthreads();
  const float arg = inv > 0.0f ? bmax / inv : 0.0f;
  if (arg <= 1.5f) return 13;
  if (arg <= 2.8f) return 17;
  if (arg <= 4.3f) return 21;
  if (arg <= 5.9f) return 25;
  return kChebMoments;
}

// log v! outside the shared-memory table (v >= 64, or data outside the support): kept out of line
__device__ __noinline__ float poisson_log_factorial_slow(float v) { return log_factorial(v); }

// per-CTA output row: mu_0 .. mu_32, V = sum v, Vx = sum v x, C = sum log v!, n  (live elements)
//
// Four elements per thread and iteration (one 128-bit load each of x and v, four mask bytes) give
// four independent recurrence chains; the next iteration's loads are issued before the current
// one is consumed, and the loop is instantiated for each moment count of cheb_moments_needed.
// Packed FFMA2 / FADD2 were tried and rejected: `tools/f32x2_probe.cu` measures the same 128
// fp32 lanes per clock and SM for packed and scalar code, and the packed kernel's register pairs
// cost a third of the occupancy (445 us against 394 us at N = 1e8).
__global__ void __launch_bounds__(kChebThreads, 3)
poisson_moment_kernel(mnf_site_t st, const float* __restrict__ range_partial, int n_range,
                      const RangeSlot* __restrict__ slot, const float* __restrict__ z, int S, int D,
                      double* __restrict__ cta_out, float* __restrict__ tmax_out, uint32_t* __restrict__ status) {
  extern __shared__ double s_mom[];                     // [kChebMoments + 3][kChebThreads] fp64 running sums
  __shared__ float s_tmp[2 * kChebThreads / 32];
  __shared__ float s_logfact[64];
  __shared__ double s_red[kChebThreads / 32];
  const int64_t n = st.numel;
  const float* __restrict__ value = st.value;
  const float* __restrict__ xp = st.param[0].x;
  const uint8_t* __restrict__ mask = st.mask;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < 64) s_logfact[threadIdx.x] = kLogFactorial[threadIdx.x];
  const ChebRange range = cheb_range(range_partial, n_range, s_tmp, slot);   // has the block barriers
  const float mid = range.mid, inv = range.inv;
  float tmax = 0.0f;                                    // max |t| over the live elements: verifies a cached range
  const int n_moments = cheb_moments_needed(st.param[0], z, S, D, inv, s_tmp);
#pragma unroll
  for (int j = 0; j < kChebMoments + 3; ++j) s_mom[j * kChebThreads + threadIdx.x] = 0.0;

  float mu[kChebMoments];
#pragma unroll
  for (int j = 0; j < kChebMoments; ++j) mu[j] = 0.0f;
  float f_v = 0.0f, f_vx = 0.0f, f_c = 0.0f;
  unsigned int cnt = 0;
  bool bad_value = false;
  constexpr int kFlushEvery = 32;                       // 128 elements per fp32 running sum
  const int64_t tid = (int64_t)blockIdx.x * kChebThreads + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * kChebThreads;
  const int64_t n4 = n >> 2;

  auto run = [&](auto nm_tag) {
    constexpr int NM = decltype(nm_tag)::value;
    auto flush = [&]() {
#pragma unroll
      for (int j = 0; j < NM; ++j) {
        s_mom[j * kChebThreads + threadIdx.x] += (double)mu[j];
        mu[j] = 0.0f;
      }
      s_mom[(kChebMoments + 0) * kChebThreads + threadIdx.x] += (double)f_v;
      s_mom[(kChebMoments + 1) * kChebThreads + threadIdx.x] += (double)f_vx;
      s_mom[(kChebMoments + 2) * kChebThreads + threadIdx.x] += (double)f_c;
      f_v = 0.0f; f_vx = 0.0f; f_c = 0.0f;
    };
    // dead elements enter with weight 0: the recurrence is linear in (T_0, T_1), so every T_j is 0
    auto take4 = [&](const float4& x4, const float4& v4, uint32_t m4) {
      const float x[4] = {x4.x, x4.y, x4.z, x4.w}, v[4] = {v4.x, v4.y, v4.z, v4.w};
      float w0[4], w1[4], tp[4], cv[4];
      bool all_small = true;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const bool live = ((m4 >> (8 * e)) & 0xffu) != 0u;
        const float xs = live ? x[e] : 0.0f;
        cv[e] = live ? v[e] : 0.0f;
        const float t = live ? (x[e] - mid) * inv : 0.0f;
        w0[e] = live ? 1.0f : 0.0f;
        w1[e] = t;
        tmax = fmaxf(tmax, fabsf(t));
        tp[e] = t + t;
        cnt += live ? 1u : 0u;
        f_v += cv[e];
        f_vx = fmaf(cv[e], xs, f_vx);
        // integer in [0, 64): 2^23 + cv is exact and carries cv in its low mantissa bits
        const float big = cv[e] + 8388608.0f;
        const bool small = cv[e] >= 0.0f && cv[e] < 64.0f && (big - 8388608.0f) == cv[e];
        f_c += small ? s_logfact[__float_as_uint(big) & 63u] : 0.0f;
        all_small = all_small && small;
      }
      if (!all_small) {                                 // rare: large counts, or data outside the support
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float big = cv[e] + 8388608.0f;
          if (!(cv[e] >= 0.0f && cv[e] < 64.0f && (big - 8388608.0f) == cv[e])) {
            if (!in_support(MNF_POISSON, cv[e])) bad_value = true;
            f_c += poisson_log_factorial_slow(cv[e]);
          }
        }
      }
      mu[0] += (w0[0] + w0[1]) + (w0[2] + w0[3]);
      mu[1] += (w1[0] + w1[1]) + (w1[2] + w1[3]);
#pragma unroll
      for (int j = 2; j < NM; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float wn = fmaf(tp[e], w1[e], -w0[e]);
          w0[e] = w1[e];
          w1[e] = wn;
        }
        mu[j] += (w1[0] + w1[1]) + (w1[2] + w1[3]);
      }
    };
    auto fetch = [&](int64_t g, float4& x4, float4& v4, uint32_t& m4) {
      if (g < n4) {
        x4 = __ldg(reinterpret_cast<const float4*>(xp) + g);
        v4 = __ldg(reinterpret_cast<const float4*>(value) + g);
        m4 = mask != nullptr ? __ldg(reinterpret_cast<const uint32_t*>(mask) + g) : 0x01010101u;
      } else {
        m4 = 0u;
      }
    };
    float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f), v4 = x4, nx4 = x4, nv4 = x4;
    uint32_t m4 = 0u, nm4 = 0u;
    int pending = 0;
    fetch(tid, x4, v4, m4);
    for (int64_t g = tid; g < n4; g += nth) {
      fetch(g + nth, nx4, nv4, nm4);             // in flight while this group is consumed
      take4(x4, v4, m4);
      if (++pending == kFlushEvery) { flush(); pending = 0; }
      x4 = nx4; v4 = nv4; m4 = nm4;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {   // ragged tail: at most three elements, one warp
      const int64_t i = (n4 << 2) + threadIdx.x;
      const bool inb = i < n;
      const float4 tx = make_float4(inb ? __ldg(xp + i) : 0.0f, 0.0f, 0.0f, 0.0f);
      const float4 tv = make_float4(inb ? __ldg(value + i) : 0.0f, 0.0f, 0.0f, 0.0f);
      take4(tx, tv, (inb && (mask == nullptr || __ldg(mask + i) != 0)) ? 1u : 0u);
    }
    flush();
  };
  switch (n_moments) {
    case 13: run(std::integral_constant<int, 13>{}); break;
    case 17: run(std::integral_constant<int, 17>{}); break;
    case 21: run(std::integral_constant<int, 21>{}); break;
    case 25: run(std::integral_constant<int, 25>{}); break;
    default: run(std::integral_constant<int, kChebMoments>{}); break;
  }
  __syncthreads();
  // column j: warp (j mod 8) adds the 256 thread slots in fixed order
  double* out = cta_out + (size_t)blockIdx.x * kChebCols;
  for (int j = warp; j < kChebMoments + 3; j += kChebThreads / 32) {
    double t = 0.0;
#pragma unroll
    for (int k = 0; k < kChebThreads / 32; ++k) t += s_mom[j * kChebThreads + k * 32 + lane];
    t = warp_sum(t);
    if (lane == 0) out[j] = t;
  }
  const double c = warp_sum((double)cnt);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
  if (lane == 0) { s_red[warp] = c; s_tmp[warp] = tmax; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kChebThreads / 32; ++w) { t += s_red[w]; tmax = fmaxf(tmax, s_tmp[w]); }
    out[kChebMoments + 3] = t;
    tmax_out[blockIdx.x] = tmax == tmax ? tmax : INFINITY;    // a NaN covariate must not pass the check
  }
  if (bad_value) atomicOr(status, MNF_ST_BAD_VALUE);
}

inline size_t poisson_moment_smem_bytes() { return sizeof(double) * (kChebMoments + 3) * kChebThreads; }

// modified Bessel functions I_0(b) .. I_{jmax}(b), 0 <= b <= kChebMaxArg: Miller's downward
// recurrence p_{j-1} = p_{j+1} + (2j / b) p_j from j = jmax + 24, normalised with
// exp(b) = I_0 + 2 sum_{j>=1} I_j; below b = 1e-2 the leading terms of the power series.
__device__ inline void bessel_i(double b, int jmax, double* out) {
  if (b < 1e-2) {
    const double h = 0.5 * b, q = h * h;
    double pref = 1.0;
    for (int j = 0; j <= jmax; ++j) {
      if (j > 0) pref *= h / (double)j;
      out[j] = pref * (1.0 + q / (double)(j + 1) * (1.0 + q / (2.0 * (double)(j + 2))));
    }
    return;
  }
  const double two_over_b = 2.0 / b;
  double p_hi = 0.0, p = 1e-250, sum = 0.0;     // (2j/b)^(jmax+24) <= 1e4^57: no overflow
  for (int j = jmax + 24; j >= 1; --j) {
    const double p_lo = p_hi + (double)j * two_over_b * p;
    sum += p;                                   // p = p_j
    if (j <= jmax) out[j] = p;
    p_hi = p;
    p = p_lo;
  }
  out[0] = p;
  const double norm = exp(b) / (p + 2.0 * sum);
  for (int j = 0; j <= jmax; ++j) out[j] *= norm;
}

// One block: sum the CTA rows in fixed order, evaluate the expansion per particle, check it, and
// either add the site to the step accumulator (need_exact = 0) or leave it to the exact kernel.
__global__ void __launch_bounds__(kChebThreads)
poisson_moment_finish_kernel(mnf_site_t st, const double* __restrict__ cta_rows, int n_cta,
                             const float* __restrict__ range_partial, int n_range, RangeSlot* __restrict__ slot,
                             const float* __restrict__ tmax_partial,
                             const float* __restrict__ z, int S, int D, double* __restrict__ acc,
                             uint32_t* __restrict__ need_exact) {
  __shared__ double s_tot[kChebCols];
  __shared__ float s_tmp[2 * kChebThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const ChebRange range = cheb_range(range_partial, n_range, s_tmp, slot);
  // the largest |t| the moment pass saw: 1 (up to rounding) when the range is this step's own
  float tmax = 0.0f;
  for (int b = threadIdx.x; b < n_cta; b += blockDim.x) tmax = fmaxf(tmax, tmax_partial[b]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
  if (lane == 0) s_tmp[warp] = tmax;
  __syncthreads();
  for (int w = 0; w < kChebThreads / 32; ++w) tmax = fmaxf(tmax, s_tmp[w]);
  __syncthreads();
  const bool range_ok = tmax <= 1.0f + 1e-5f;
  const int nm = cheb_moments_needed(st.param[0], z, S, D, range.inv, s_tmp);     // moments the sweep kept
  for (int c = warp; c < kChebCols; c += kChebThreads / 32) {
    double t = 0.0;
    for (int b = lane; b < n_cta; b += 32) t += cta_rows[(size_t)b * kChebCols + c];
    t = warp_sum(t);
    if (lane == 0) s_tot[c] = t;
  }
  __syncthreads();
  const double V = s_tot[kChebMoments], Vx = s_tot[kChebMoments + 1], C = s_tot[kChebMoments + 2];
  const double cnt = s_tot[kChebMoments + 3];
  const double mid = (double)range.mid, half = range.inv > 0.0f ? 1.0 / (double)range.inv : 0.0;
  const mnf_link_t L0 = st.param[0];

  bool ok = true;
  double lp = 0.0, dA = 0.0, dB = 0.0;
  const int s = threadIdx.x;
  if (s < S) {
    const float* zs = z + (int64_t)s * D;
    const double a_s = (double)(L0.a_const + (L0.a_lat >= 0 ? zs[L0.a_lat] : 0.0f));
    const double b_s = (double)(L0.b_const + (L0.b_lat >= 0 ? zs[L0.b_lat] : 0.0f));
    const double arg = b_s * half;
    ok = fabs(arg) <= kChebMaxArg && isfinite(a_s);          // false for NaN / inf too
    if (ok) {
      double f[kChebMoments + 1];                            // f_j = sign^j I_j(|arg|), j = 0 .. J + 1
      bessel_i(fabs(arg), kChebMoments, f);
      if (arg < 0.0)
        for (int j = 1; j <= kChebMoments; j += 2) f[j] = -f[j];
      double s0 = f[0] * s_tot[0], mag = fabs(s0);
      double s1 = f[1] * s_tot[0];
      for (int j = 1; j < nm; ++j) {
        const double term = 2.0 * f[j] * s_tot[j];
        s0 += term;
        mag += fabs(term);
        s1 += (f[j - 1] + f[j + 1]) * s_tot[j];
      }
      // |T_j| <= 1: the dropped tail is at most 2 n sum_{j > J} I_j, dominated by its first term
      const double tail = 4.0 * fabs(f[nm]) * cnt;
      ok = cnt == 0.0 || (s0 > 0.0 && mag <= kChebMaxCancel * s0 && tail <= kChebMaxTrunc * s0);
      const double base = exp(a_s + b_s * mid);
      const double R = base * s0, Rx = mid * R + half * base * s1;
      ok = ok && isfinite(R) && isfinite(Rx);
      lp = a_s * V + b_s * Vx - R - C;
      dA = V - R;
      dB = Vx - Rx;
    }
  }
  const int any_bad = __syncthreads_or(ok && range_ok ? 0 : 1);
  if (threadIdx.x == 0) {
    *need_exact = any_bad ? 1u : 0u;
    if (slot != nullptr) {
      if (!range.cached) {            // this step ran the range pass: keep its result for the next steps
        slot->lo = range.lo;
        slot->hi = range.hi;
        slot->valid = range.lo <= range.hi ? 1u : 0u;
      } else if (any_bad || (tmax < 0.5f && range.inv > 0.0f)) {
        slot->valid = 0u;             // the data left the range (or use little of it): measure it again next step
      }
    }
  }
  if (any_bad || s >= S) return;
  const double w = st.scale;
  double* as = acc + (int64_t)s * (D + 1);
  atomicAdd(as, w * lp);
  if (L0.a_lat >= 0) atomicAdd(as + 1 + L0.a_lat, w * dA);
  if (L0.b_lat >= 0) atomicAdd(as + 1 + L0.b_lat, w * dB);
}

// ---- Normal(A_s + B_s x, sigma_s): data-only sufficient statistics -------------------------------
constexpr int kStatThreads = 256;
constexpr int kStatCols = 6;   // n, Sx, Sxx, Sv, Svx, Svv

struct NormalStats {
  double sx = 0.0, sxx = 0.0, sv = 0.0, svx = 0.0, svv = 0.0;
  unsigned int n = 0;
  bool bad = false;
  __device__ __forceinline__ void add(float v, float x, bool live) {
    if (live) {
      const double dv = (double)v, dx = (double)x;
      ++n;
      sx += dx;
      sxx = fma(dx, dx, sxx);
      sv += dv;
      svx = fma(dv, dx, svx);
      svv = fma(dv, dv, svv);
      bad |= v != v;
    }
  }
};

// VEC: value (and x, mask) are 16-byte (4-byte for the mask) aligned with unit stride: 128-bit loads
template <bool VEC>
__global__ void __launch_bounds__(kStatThreads)
normal_stats_kernel(mnf_site_t st, double* __restrict__ cta_stats, uint32_t* __restrict__ status) {
  const int64_t n = st.numel;
  const float* __restrict__ value = st.value;
  const float* __restrict__ xp = st.param[0].x;
  const int64_t xs = st.param[0].x_stride;
  const uint8_t* __restrict__ mask = st.mask;
  const int64_t tid = (int64_t)blockIdx.x * kStatThreads + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * kStatThreads;
  NormalStats acc;
  if (VEC) {
    const int64_t n4 = n >> 2;
    for (int64_t g = tid; g < n4; g += nth) {
      const float4 v4 = __ldg(reinterpret_cast<const float4*>(value) + g);
      const float4 x4 = xp != nullptr ? __ldg(reinterpret_cast<const float4*>(xp) + g) : make_float4(1.f, 1.f, 1.f, 1.f);
      const uint32_t m4 = mask != nullptr ? __ldg(reinterpret_cast<const uint32_t*>(mask) + g) : 0x01010101u;
      acc.add(v4.x, x4.x, (m4 & 0x000000ffu) != 0);
      acc.add(v4.y, x4.y, (m4 & 0x0000ff00u) != 0);
      acc.add(v4.z, x4.z, (m4 & 0x00ff0000u) != 0);
      acc.add(v4.w, x4.w, (m4 & 0xff000000u) != 0);
    }
    const int64_t i = (n4 << 2) + tid;     // ragged tail: at most three elements
    if (i < n) acc.add(__ldg(value + i), xp != nullptr ? __ldg(xp + i) : 1.0f, mask == nullptr || __ldg(mask + i) != 0);
  } else {
    for (int64_t i = tid; i < n; i += nth)
      acc.add(__ldg(value + i), xp != nullptr ? __ldg(xp + xs * i) : 1.0f, mask == nullptr || __ldg(mask + i) != 0);
  }
  __shared__ double s_red[kStatThreads / 32][kStatCols];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double c[kStatCols] = {(double)acc.n, acc.sx, acc.sxx, acc.sv, acc.svx, acc.svv};
#pragma unroll
  for (int k = 0; k < kStatCols; ++k) c[k] = warp_sum(c[k]);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < kStatCols; ++k) s_red[warp][k] = c[k];
  __syncthreads();
  if (threadIdx.x < kStatCols) {
    double t = 0.0;
    for (int w = 0; w < kStatThreads / 32; ++w) t += s_red[w][threadIdx.x];
    cta_stats[(size_t)blockIdx.x * kStatCols + threadIdx.x] = t;
  }
  if (acc.bad) atomicOr(status, MNF_ST_BAD_VALUE);
}

// One block of kStatCols warps: warp k sums statistic k over the CTAs in fixed order; then one
// thread per particle evaluates the closed forms and adds them to the step accumulator.
__global__ void __launch_bounds__(32 * kStatCols)
normal_stats_finish_kernel(mnf_site_t st, const double* __restrict__ cta_stats, int n_cta,
                           const float* __restrict__ z, int S, int D, double* __restrict__ acc,
                           uint32_t* __restrict__ status) {
  __shared__ double s_tot[kStatCols];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double t = 0.0;
  for (int b = lane; b < n_cta; b += 32) t += cta_stats[(size_t)b * kStatCols + warp];
  t = warp_sum(t);
  if (lane == 0) s_tot[warp] = t;
  __syncthreads();
  const double cnt = s_tot[0], Sx = s_tot[1], Sxx = s_tot[2], Sv = s_tot[3], Svx = s_tot[4], Svv = s_tot[5];
  const mnf_link_t L0 = st.param[0], L1 = st.param[1];
  bool bad_param = false;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const float* zs = z + (int64_t)s * D;
    const double A = (double)(L0.a_const + (L0.a_lat >= 0 ? zs[L0.a_lat] : 0.0f));
    const double B = (double)(L0.b_const + (L0.b_lat >= 0 ? zs[L0.b_lat] : 0.0f));
    const float u = (L1.a_const + (L1.a_lat >= 0 ? zs[L1.a_lat] : 0.0f)) +
                    (L1.b_const + (L1.b_lat >= 0 ? zs[L1.b_lat] : 0.0f));    // x == NULL means x == 1
    const float sigma_f = L1.transform == MNF_T_EXP ? expf(u) : u;
    if (!(sigma_f > 0.0f)) bad_param = true;
    const double T1 = Sv - A * cnt - B * Sx;
    const double Tx = Svx - A * Sx - B * Sxx;
    const double T2 = Svv - A * Sv - B * Svx - A * T1 - B * Tx;
    const double sigma = (double)sigma_f, inv = 1.0 / sigma, iv = inv * inv;
    const double w = st.scale;
    const double lp = -0.5 * iv * T2 - cnt * (log(sigma) + 0.91893853320467274178);
    const double d1 = (iv * T2 - cnt) * inv * (L1.transform == MNF_T_EXP ? sigma : 1.0);
    double* as = acc + (int64_t)s * (D + 1);
    // several links may name the same latent column, hence atomics (launches are stream-ordered)
    atomicAdd(as, w * lp);
    if (L0.a_lat >= 0) atomicAdd(as + 1 + L0.a_lat, w * iv * T1);
    if (L0.b_lat >= 0) atomicAdd(as + 1 + L0.b_lat, w * iv * Tx);
    if (L1.a_lat >= 0) atomicAdd(as + 1 + L1.a_lat, w * d1);
    if (L1.b_lat >= 0) atomicAdd(as + 1 + L1.b_lat, w * d1);
  }
  if (bad_param) atomicOr(status, MNF_ST_BAD_PARAM);
}

}  // namespace mnf
