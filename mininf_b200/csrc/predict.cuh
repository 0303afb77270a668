// Batched posterior-predictive sampler: mnf_predictive of include/mininf_b200.h.
//
// Replaces the per-sample Python loop of `broadcast_samples` (mininf/core.py:548-584: for every
// posterior sample the model is re-run under a SampleTracer, mininf/core.py:192-204) with ONE
// launch: the host traces the model once, and block b of the grid walks the traced sites in model
// order for posterior sample b - deterministic `value` sites (mininf/core.py:390-492) are
// evaluated from their link expression, sites the samples do not provide are drawn from their
// distribution (torch's samplers: Normal loc + scale eps, Gamma by Marsaglia-Tsang, Beta as a ratio
// of gammas, Bernoulli by inversion, Poisson by Knuth's product / Hoermann's PTRS as in ATen's
// Distributions.h:147-206) - writing every result into row b of z, where later sites read it.
// Included by abi.cu.
#pragma once

#include "common.cuh"
#include "small.cuh"

namespace mnf {

__device__ __forceinline__ float philox_uniform(uint32_t w) {   // (0, 1]
  return ((float)(w >> 8) + 1.0f) * (1.0f / 16777216.0f);
}

// Poisson(rate) draw on a private Philox stream
__device__ inline float philox_poisson(float rate, uint64_t seed, uint64_t offset, uint64_t index) {
  if (!(rate > 0.0f)) return 0.0f;
  uint32_t round = 0;
  if (rate < 10.0f) {
    // Knuth: count uniforms until their product drops below exp(-rate)
    const float limit = expf(-rate);
    float prod = 1.0f;
    int k = 0;
    for (; round < 64; ++round) {
      Philox rng(seed, offset, kPhiloxPredict | (index << 8) | round);
      const uint4 r = rng.next();
      const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        prod *= philox_uniform(w[q]);
        if (prod <= limit) return (float)k;
        ++k;
      }
    }
    return (float)k;
  }
  // PTRS (Hoermann 1993), the transformed-rejection sampler ATen uses for rate >= 10
  const float slam = sqrtf(rate), loglam = logf(rate);
  const float b = 0.931f + 2.53f * slam;
  const float a = -0.059f + 0.02483f * b;
  const float invalpha = 1.1239f + 1.1328f / (b - 3.4f);
  const float vr = 0.9277f - 3.6224f / (b - 2.0f);
  for (; round < 64; ++round) {
    Philox rng(seed, offset, kPhiloxPredict | (index << 8) | round);
    const uint4 r = rng.next();
    const float U = philox_uniform(r.x) - 0.5f, V = philox_uniform(r.y);
    const float us = 0.5f - fabsf(U);
    const float k = floorf((2.0f * a / us + b) * U + rate + 0.43f);
    if (us >= 0.07f && V <= vr) return k;
    if (k < 0.0f || (us < 0.013f && V > us)) continue;
    if (logf(V) + logf(invalpha) - logf(a / (us * us) + b) <= -rate + k * loglam - lgammaf(k + 1.0f)) return k;
  }
  return floorf(rate);
}

// one draw from `family` with (already transformed) parameters p0, p1
__device__ inline float predictive_draw(int family, float p0, float p1, uint64_t seed, uint64_t offset, uint64_t index,
                                        uint32_t& bad) {
  switch (family) {
    case MNF_NORMAL: {
      if (!(p1 > 0.0f)) bad |= MNF_ST_BAD_PARAM;
      Philox rng(seed, offset, kPhiloxPredict | (index << 8));
      const uint4 r = rng.next();
      return fmaf(box_muller(r.x, r.y).x, p1, p0);
    }
    case MNF_GAMMA:
      if (!(p0 > 0.0f) || !(p1 > 0.0f)) bad |= MNF_ST_BAD_PARAM;
      return fmaxf(philox_standard_gamma(p0, seed, offset, kPhiloxPredictGamma | index) / p1, kFloatTiny);
    case MNF_BETA: {
      if (!(p0 > 0.0f) || !(p1 > 0.0f)) bad |= MNF_ST_BAD_PARAM;
      const float g1 = philox_standard_gamma(p0, seed, offset, kPhiloxPredictGamma | (2 * index));
      const float g0 = philox_standard_gamma(p1, seed, offset, kPhiloxPredictGamma | (2 * index + 1));
      return fminf(fmaxf(g1 / (g1 + g0), kFloatEps), 1.0f - kFloatEps);
    }
    case MNF_BERNOULLI_PROBS:
    case MNF_BERNOULLI_LOGITS: {
      const float prob = family == MNF_BERNOULLI_LOGITS ? sigmoid_f(p0) : p0;
      if (!(prob >= 0.0f && prob <= 1.0f)) bad |= MNF_ST_BAD_PARAM;
      Philox rng(seed, offset, kPhiloxPredict | (index << 8));
      const uint4 r = rng.next();
      return ((float)(r.x >> 8) * (1.0f / 16777216.0f)) < prob ? 1.0f : 0.0f;      // uniform in [0, 1)
    }
    case MNF_POISSON:
      if (!(p0 >= 0.0f)) bad |= MNF_ST_BAD_PARAM;
      return philox_poisson(p0, seed, offset, index);
    default:
      bad |= MNF_ST_BAD_PARAM;
      return 0.0f;
  }
}

constexpr int kPredictThreads = 256;

// grid = posterior samples; z [B][n_columns]: given samples in their columns, results written in place
__global__ void __launch_bounds__(kPredictThreads)
predictive_kernel(const mnf_pred_site_t* __restrict__ sites, int n_sites, int n_columns, float* __restrict__ z,
                  uint64_t seed, uint64_t offset, uint32_t* __restrict__ status) {
  const int b = blockIdx.x;
  float* zb = z + (int64_t)b * n_columns;
  uint32_t bad = 0;
  for (int k = 0; k < n_sites; ++k) {
    const mnf_pred_site_t site = sites[k];
    const bool two = site.family <= MNF_BETA;
    for (int64_t i = threadIdx.x; i < site.numel; i += kPredictThreads) {
      float p0;
      if (site.X != nullptr) {
        // dense link: T(icpt + X[i, :] . theta) with theta a block of this sample's columns
        float eta = site.icpt_const + (site.icpt_lat >= 0 ? zb[site.icpt_lat] : 0.0f);
        const float* row = site.X + i * site.ldx;
        for (int j = 0; j < site.p; ++j) eta = fmaf(__ldg(row + j), zb[site.theta_lat + j], eta);
        p0 = site.transform == MNF_T_EXP ? expf(eta) : eta;
      } else {
        p0 = eval_link(site.param[0], zb, i).value;
      }
      const float p1 = two ? eval_link(site.param[1], zb, i).value : 0.0f;
      const uint64_t index = (uint64_t)b * (uint64_t)n_columns + (uint64_t)(site.out_col + i);
      zb[site.out_col + i] = site.kind == MNF_PRED_VALUE ? p0 : predictive_draw(site.family, p0, p1, seed, offset, index, bad);
    }
    __syncthreads();      // later sites of this sample read what this one wrote
  }
  if (bad) atomicOr(status, bad);
}

}  // namespace mnf
