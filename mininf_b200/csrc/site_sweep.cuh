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
// Specialised single-site sweeps for the two hot site kinds of the missing-observations config.
// The generic kernel above spends ~40 instructions per (element, particle); here everything that
// does not depend on the particle is summed once per element, and the per-particle work shrinks to
// the irreducible running sums:
//   Poisson(exp(A_s + B_s x)):   R_s = sum rate, Rx_s = sum rate*x          (5 instructions)
//       log p = A_s V + B_s Vx - R_s - C,  d/dA = V - R_s,  d/dB = Vx - Rx_s
//       with V = sum v, Vx = sum v*x, C = sum log v! over the live elements
//   Normal(A_s + B_s x, sigma_s): T1 = sum r, Tx = sum r*x, T2 = sum r^2, r = v - A_s - B_s x (5)
//       log p = -T2 / (2 sigma^2) - n log(sigma sqrt(2 pi)), d/dA = T1/sigma^2, d/dB = Tx/sigma^2,
//       d/dsigma = (T2/sigma^2 - n)/sigma
// Same mapping as the generic kernel (particles on lanes, elements broadcast from a per-warp
// staging slot, masked elements skipped warp-uniformly) and the same partial layout as its
// one-site instance: [S][5] = weight * (log p, du0, du0*x0, du1, du1*x1).
// ---------------------------------------------------------------------------------------------
constexpr int kFastNone = -1, kFastPoissonExp = 0, kFastNormalId = 1;

inline int site_fast_kind(const mnf_site_t& st) {
  if (st.family == MNF_POISSON && st.param[0].transform == MNF_T_EXP) return kFastPoissonExp;
  if (st.family == MNF_NORMAL && st.param[0].transform == MNF_T_ID && st.param[1].x == nullptr)
    return kFastNormalId;
  return kFastNone;
}

// exp(u) as ex2(u * log2(e)) with the product carried in two terms (relative error ~2 ulp)
__device__ __forceinline__ float fast_exp(float u) {
  const float t = fmaf(u, 1.925963033500011e-8f, u * 1.4426950216293335f);
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
  return r;
}

template <int KIND, int Q>
__global__ void __launch_bounds__(kSweepThreads, 3)
site_fast_kernel(mnf_site_t st, const float* __restrict__ z, int S, int D, float* __restrict__ partial,
                 uint32_t* __restrict__ status) {
  constexpr int NA = KIND == kFastPoissonExp ? 2 : 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n = st.numel;
  const mnf_link_t L0 = st.param[0], L1 = st.param[1];

  float A[Q], B[Q];
#pragma unroll
  for (int q = 0; q < Q; ++q) {
    const int s = lane + 32 * q;
    const float* zs = z + (int64_t)(s < S ? s : 0) * D;
    A[q] = L0.a_const + (L0.a_lat >= 0 ? zs[L0.a_lat] : 0.0f);
    B[q] = L0.b_const + (L0.b_lat >= 0 ? zs[L0.b_lat] : 0.0f);
  }
  extern __shared__ double s_fast[];
  double* s_sums = s_fast;                                          // [warp][Q*32][NA]
  double* s_elem = s_fast + (size_t)kSweepWarps * Q * 32 * NA;      // [warp][4]: V, Vx, C, n
  float2* stage = reinterpret_cast<float2*>(s_elem + kSweepWarps * 4) + warp * 32;

  double run[Q][NA];
#pragma unroll
  for (int q = 0; q < Q; ++q)
#pragma unroll
    for (int a = 0; a < NA; ++a) run[q][a] = 0.0;
  double e_v = 0.0, e_vx = 0.0, e_c = 0.0;     // this lane's share of the per-element sums
  int e_n = 0;
  bool bad_value = false;

  const int64_t n_chunks = (n + 31) / 32;
  const int64_t warp_global = (int64_t)blockIdx.x * kSweepWarps + warp;
  const int64_t warps_total = (int64_t)gridDim.x * kSweepWarps;
  for (int64_t chunk = warp_global; chunk < n_chunks; chunk += warps_total) {
    const int64_t i = chunk * 32 + lane;
    const bool inb = i < n;
    const bool live = inb && (st.mask == nullptr || st.mask[i] != 0);
    const float v = live ? __ldg(st.value + i) : 0.0f;
    const float x = (inb && L0.x != nullptr) ? __ldg(L0.x + (int64_t)L0.x_stride * i) : 1.0f;
    if (live) {
      if (!in_support(st.family, v)) bad_value = true;
      ++e_n;
      if (KIND == kFastPoissonExp) {
        e_v += (double)v;
        e_vx += (double)(v * x);
        e_c += (double)log_factorial(v);
      }
    }
    const uint32_t live_bits = __ballot_sync(0xffffffffu, live);
    stage[lane] = make_float2(v, x);
    __syncwarp();
    float part[NA][Q];
#pragma unroll
    for (int a = 0; a < NA; ++a)
#pragma unroll
      for (int q = 0; q < Q; ++q) part[a][q] = 0.0f;
#pragma unroll 4
    for (int e = 0; e < 32; ++e) {
      if (!((live_bits >> e) & 1u)) continue;   // warp-uniform
      const float2 el = stage[e];
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        if (KIND == kFastPoissonExp) {
          const float rate = fast_exp(fmaf(B[q], el.y, A[q]));
          part[0][q] += rate;
          part[1][q] = fmaf(rate, el.y, part[1][q]);
        } else {
          const float r = fmaf(-B[q], el.y, el.x - A[q]);
          part[0][q] += r;
          part[1][q] = fmaf(r, el.y, part[1][q]);
          part[2][q] = fmaf(r, r, part[2][q]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int a = 0; a < NA; ++a) run[q][a] += (double)part[a][q];
    __syncwarp();   // the staging slot is rewritten by the next chunk
  }

#pragma unroll
  for (int q = 0; q < Q; ++q)
#pragma unroll
    for (int a = 0; a < NA; ++a) s_sums[((size_t)warp * Q * 32 + q * 32 + lane) * NA + a] = run[q][a];
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

  // one particle per thread: combine the warps in fixed order and finish the closed forms
  bool bad_param = false;
  for (int s = threadIdx.x; s < S; s += kSweepThreads) {
    double t[NA], el[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int a = 0; a < NA; ++a) t[a] = 0.0;
    for (int w = 0; w < kSweepWarps; ++w) {
#pragma unroll
      for (int a = 0; a < NA; ++a) t[a] += s_sums[((size_t)w * Q * 32 + s) * NA + a];
#pragma unroll
      for (int a = 0; a < 4; ++a) el[a] += s_elem[w * 4 + a];
    }
    const float* zs = z + (int64_t)s * D;
    const double a_s = (double)(L0.a_const + (L0.a_lat >= 0 ? zs[L0.a_lat] : 0.0f));
    const double b_s = (double)(L0.b_const + (L0.b_lat >= 0 ? zs[L0.b_lat] : 0.0f));
    double lp, d0, d0x, d1 = 0.0;
    if (KIND == kFastPoissonExp) {
      lp = a_s * el[0] + b_s * el[1] - t[0] - el[2];
      d0 = el[0] - t[0];
      d0x = el[1] - t[1];
    } else {
      const float u = (L1.a_const + (L1.a_lat >= 0 ? zs[L1.a_lat] : 0.0f)) +
                      (L1.b_const + (L1.b_lat >= 0 ? zs[L1.b_lat] : 0.0f));    // x == NULL means x == 1
      const float sigma_f = L1.transform == MNF_T_EXP ? expf(u) : u;
      if (!(sigma_f > 0.0f)) bad_param = true;
      const double sigma = (double)sigma_f, inv = 1.0 / sigma, iv = inv * inv;
      lp = -0.5 * iv * t[2] - el[3] * ((double)logf(sigma_f) + (double)kLogSqrt2Pi);
      d0 = iv * t[0];
      d0x = iv * t[1];
      d1 = (iv * t[2] - el[3]) * inv * (L1.transform == MNF_T_EXP ? sigma : 1.0);
    }
    float* out = partial + ((size_t)blockIdx.x * S + s) * 5;
    const double w = st.scale;
    out[0] = (float)(w * lp);
    out[1] = (float)(w * d0);
    out[2] = (float)(w * d0x);
    out[3] = (float)(w * d1);
    out[4] = (float)(w * d1);
  }
  if (bad_value) atomicOr(status, MNF_ST_BAD_VALUE);
  if (bad_param) atomicOr(status, MNF_ST_BAD_PARAM);
}

template <int KIND, int Q>
inline size_t site_fast_smem_bytes() {
  constexpr int NA = KIND == kFastPoissonExp ? 2 : 3;
  return sizeof(double) * kSweepWarps * Q * 32 * NA + sizeof(double) * kSweepWarps * 4 +
         sizeof(float2) * kSweepWarps * 32;
}

}  // namespace mnf
