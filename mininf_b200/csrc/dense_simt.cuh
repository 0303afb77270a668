// Dense-link sweep, fp32 SIMT variant (MNF_DENSE_FP32): exact fp32 FMA arithmetic for any feature
// count p and particle count S. It is the precision reference of the tcgen05 variant and the path
// for shapes the tensor-core kernel does not cover. One pass over X and y serves all particles
// and both directions:
//   phase 1  eta[r][s] = a_s + sum_j X[r][j] theta[s][j];  log-density and score R[r][s]
//   phase 2  G[j][s]  += sum_r X[r][j] R[r][s]
// (replaces aten::mv/addmv_, MvBackward and the elementwise log_prob chain of the reference,
// SURVEY.md §2.1; user link code tests/test_mininf.py:11, examples/minibatch.md:33).
#pragma once

#include "common.cuh"

namespace mnf {

constexpr int kSimtThreads = 256;
constexpr int kSimtRows = 32;  // rows per tile

// per-particle parameters of the dense site, staged once per CTA
struct DenseParticle {
  float icpt;     // intercept a_s
  float scale;    // Normal scale sigma_s (after transform)
  float dscale;   // d sigma / d u (exp link: sigma, identity: 1)
};

__device__ __forceinline__ DenseParticle dense_particle(const mnf_dense_site_t& site,
                                                        const float* zs) {
  DenseParticle o;
  o.icpt = site.icpt_const + (site.icpt_lat >= 0 ? zs[site.icpt_lat] : 0.0f);
  o.scale = 1.0f;
  o.dscale = 0.0f;
  if (site.family == MNF_NORMAL) {
    const LinkVal l = eval_link(site.scale, zs, 0);
    o.scale = l.value;
    o.dscale = l.du;
  }
  return o;
}

// log-density and score w.r.t. eta (and w.r.t. the Normal scale) of one (row, particle) pair
__device__ __forceinline__ void dense_point(int family, float y, float eta, float sigma, float& lp,
                                            float& deta, float& dsigma) {
  dsigma = 0.0f;
  if (family == MNF_NORMAL) {
    const float inv = 1.0f / sigma;
    const float r = (y - eta) * inv;
    lp = -0.5f * r * r - logf(sigma) - kLogSqrt2Pi;
    deta = r * inv;
    dsigma = (r * r - 1.0f) * inv;
  } else if (family == MNF_BERNOULLI_LOGITS) {
    lp = y * eta - softplus_f(eta);
    deta = y - sigmoid_f(eta);
  } else {  // MNF_POISSON with the exp link: rate = exp(eta)
    const float rate = expf(eta);
    lp = y * eta - rate - lgammaf(y + 1.0f);
    deta = y - rate;
  }
}

// partial layout per CTA: [S][ncol], ncol = 1 + p + 2:
//   col 0 log-density, 1..p theta gradient, p+1 intercept gradient, p+2 scale-link gradient (du)
__global__ void __launch_bounds__(kSimtThreads)
dense_simt_kernel(mnf_dense_site_t site, const float* __restrict__ z, int S, int D,
                  float* __restrict__ partial, uint32_t* __restrict__ status) {
  extern __shared__ float smem[];
  const int p = site.p;
  const int ldt = p + 1;                       // padded row of the X tile
  double* sStat = reinterpret_cast<double*>(smem);  // [S][3] lp, d icpt, d scale-link (fp64 running sums)
  float* sTheta = reinterpret_cast<float*>(sStat + (size_t)S * 3);  // [S][p]
  float* sX = sTheta + (size_t)S * p;          // [kSimtRows][p+1]
  float* sR = sX + (size_t)kSimtRows * ldt;    // [S][kSimtRows+1] (padded: conflict-free both ways)
  float* sG = sR + (size_t)(kSimtRows + 1) * S; // [p][S] running gradient
  DenseParticle* sPar = reinterpret_cast<DenseParticle*>(sG + (size_t)p * S);  // [S]

  const int tid = threadIdx.x;
  for (int i = tid; i < S * p; i += kSimtThreads) sTheta[i] = z[(int64_t)(i / p) * D + site.theta_lat + (i % p)];
  for (int i = tid; i < p * S; i += kSimtThreads) sG[i] = 0.0f;
  for (int i = tid; i < S * 3; i += kSimtThreads) sStat[i] = 0.0;
  for (int s = tid; s < S; s += kSimtThreads) sPar[s] = dense_particle(site, z + (int64_t)s * D);
  __syncthreads();

  uint32_t bad = 0;
  if (site.family == MNF_NORMAL)
    for (int s = tid; s < S; s += kSimtThreads)
      if (!(sPar[s].scale > 0.0f)) bad |= MNF_ST_BAD_PARAM;

  const int64_t n_tiles = (site.n_rows + kSimtRows - 1) / kSimtRows;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * kSimtRows;
    // stage the X tile (coalesced along features)
    for (int i = tid; i < kSimtRows * p; i += kSimtThreads) {
      const int r = i / p, j = i % p;
      const int64_t row = row0 + r;
      sX[r * ldt + j] = row < site.n_rows ? __ldg(site.X + row * site.ldx + j) : 0.0f;
    }
    __syncthreads();
    // phase 1: one (row, particle) pair per thread iteration; threads of a warp share the
    // particle (theta broadcast) and differ in the row (conflict-free padded X rows)
    // (kSimtRows == 32 == warp size, so lane == row and the trip count is warp-uniform; a
    // particle s is owned by exactly one warp per iteration, so sStat needs no atomics)
    for (int i = tid; i < kSimtRows * S; i += kSimtThreads) {
      const int r = i % kSimtRows, s = i / kSimtRows;
      const int64_t row = row0 + r;
      const bool live = row < site.n_rows && (site.mask == nullptr || site.mask[row] != 0);
      float lp = 0.0f, deta = 0.0f, dsig = 0.0f;
      if (live) {
        float eta = sPar[s].icpt;
        const float* xr = sX + r * ldt;
        const float* th = sTheta + (size_t)s * p;
        for (int j = 0; j < p; ++j) eta = fmaf(xr[j], th[j], eta);
        const float y = __ldg(site.y + row);
        dense_point(site.family, y, eta, sPar[s].scale, lp, deta, dsig);
        if (!in_support(site.family, y)) bad |= MNF_ST_BAD_VALUE;
      }
      sR[s * (kSimtRows + 1) + r] = deta;
      lp = warp_sum(lp);
      deta = warp_sum(deta);
      dsig = warp_sum(dsig);
      if ((tid & 31) == 0) {
        sStat[s * 3 + 0] += (double)lp;
        sStat[s * 3 + 1] += (double)deta;
        sStat[s * 3 + 2] += (double)(dsig * sPar[s].dscale);
      }
    }
    __syncthreads();
    // phase 2: every thread owns fixed (j, s) cells of G; threads of a warp share j
    for (int i = tid; i < p * S; i += kSimtThreads) {
      const int s = i % S, j = i / S;
      float g = 0.0f;
#pragma unroll 8
      for (int r = 0; r < kSimtRows; ++r) g = fmaf(sX[r * ldt + j], sR[s * (kSimtRows + 1) + r], g);
      sG[i] += g;
    }
    __syncthreads();
  }

  const int ncol = 1 + p + 2;
  float* out = partial + (size_t)blockIdx.x * S * ncol;
  for (int i = tid; i < S * ncol; i += kSimtThreads) {
    const int s = i / ncol, c = i % ncol;
    float v;
    if (c == 0) v = (float)sStat[s * 3 + 0];
    else if (c <= p) v = sG[(c - 1) * S + s];
    else v = (float)sStat[s * 3 + (c - p)];
    out[i] = v;
  }
  if (bad) atomicOr(status, bad);
}

inline size_t dense_simt_smem_bytes(int S, int p) {
  return sizeof(double) * (size_t)S * 3 +
         sizeof(float) * ((size_t)S * p + (size_t)kSimtRows * (p + 1) + (size_t)(kSimtRows + 1) * S +
                          (size_t)p * S) +
         sizeof(DenseParticle) * (size_t)S;
}

}  // namespace mnf
