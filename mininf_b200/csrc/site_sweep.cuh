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

__device__ __forceinline__ bool in_support(int family, float v) {
  switch (family) {
    case MNF_NORMAL: return v == v;
    case MNF_GAMMA: return v >= 0.0f;
    case MNF_BETA: return v >= 0.0f && v <= 1.0f;
    case MNF_POISSON: return v >= 0.0f && floorf(v) == v;
    default: return v == 0.0f || v == 1.0f;
  }
}

// partial layout per CTA: [S][ncol], ncol = 1 + 4*NSITES: col 0 scaled log-density, then per site
// the scaled sums of (du0, du0*x0, du1, du1*x1)
template <int NSITES, int Q>
__global__ void __launch_bounds__(kSweepThreads)
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
  double acc[NC][Q];
#pragma unroll
  for (int c = 0; c < NC; ++c)
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[c][q] = 0.0;

  bool bad_param = false, bad_value = false;
  const int64_t n_chunks = (n + 31) / 32;
  const int64_t warp_global = (int64_t)blockIdx.x * kSweepWarps + warp;
  const int64_t warps_total = (int64_t)gridDim.x * kSweepWarps;

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
        if (st.family == MNF_POISSON) ec[k] = lgammaf(ev[k] + 1.0f);
      }
      live_bits[k] = __ballot_sync(0xffffffffu, live);
    }
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
        const float v = __shfl_sync(0xffffffffu, ev[k], e);
        const float c = __shfl_sync(0xffffffffu, ec[k], e);
        const float x0 = __shfl_sync(0xffffffffu, ex0[k], e);
        const float x1 = __shfl_sync(0xffffffffu, ex1[k], e);
        const float w = (float)st.scale;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const float u0 = fmaf(B[k][0][q], x0, A[k][0][q]);
          const float u1 = fmaf(B[k][1][q], x1, A[k][1][q]);
          float lp, du0, du1;
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
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int q = 0; q < Q; ++q) acc[c][q] += (double)part[c][q];
  }

  // cross-warp reduction in fixed order, then one partial block per CTA
  extern __shared__ double s_acc[];  // [kSweepWarps][Q*32][NC]
#pragma unroll
  for (int q = 0; q < Q; ++q)
#pragma unroll
    for (int c = 0; c < NC; ++c) s_acc[((size_t)warp * Q * 32 + q * 32 + lane) * NC + c] = acc[c][q];
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
  return sizeof(double) * kSweepWarps * Q * 32 * (1 + 4 * NSITES);
}

}  // namespace mnf
