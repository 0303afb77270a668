// O(S*D) kernels of the ELBO step: the reparameterised draw of the global latent sites, the
// element-wise evaluation of short sites (priors, small observed vectors), the fixed-order
// reduction of sweep partials into the step accumulator, and the final combine with entropy and
// pathwise gradients. None of these touches the big observed arrays.
#pragma once

#include "host.h"
#include "implicit_grad.cuh"

namespace mnf {

// Standard Gamma(alpha, 1) draw by Marsaglia & Tsang's squeeze method (ACM TOMS 26, 2000), the
// algorithm behind ATen's sample_gamma (TORCH include/ATen/native/Distributions.h:85-113), on a
// private Philox stream; alpha < 1 uses the boost Gamma(alpha + 1) * U^(1/alpha).
__device__ inline float philox_standard_gamma(float alpha, uint64_t seed, uint64_t offset, uint64_t index) {
  float scale = 1.0f;
  uint32_t round = 0;
  if (alpha < 1.0f) {
    Philox rng(seed, offset, kPhiloxGamma | (index << 8) | 0xFFu);
    const uint4 r = rng.next();
    const float u = ((float)(r.x >> 8) + 1.0f) * (1.0f / 16777216.0f);
    scale = powf(u, 1.0f / alpha);
    alpha += 1.0f;
  }
  const float d = alpha - 1.0f / 3.0f;
  const float c = rsqrtf(9.0f * d);
  for (; round < 64; ++round) {
    Philox rng(seed, offset, kPhiloxGamma | (index << 8) | round);
    const uint4 r = rng.next();
    const float x = box_muller(r.x, r.y).x;
    const float t = 1.0f + c * x;
    if (t <= 0.0f) continue;
    const float v = t * t * t;
    const float u = ((float)(r.z >> 8) + 1.0f) * (1.0f / 16777216.0f);
    const float xx = x * x;
    if (u < 1.0f - 0.0331f * xx * xx || logf(u) < 0.5f * xx + d * (1.0f - v + logf(v)))
      return fmaxf(scale * d * v, kFloatTiny);
  }
  return fmaxf(scale * d, kFloatTiny);   // unreachable in practice (acceptance > 95 % per round)
}

__device__ __forceinline__ int find_latent(const mnf_latent_t* lat, int n, int col) {
  int k = 0;
  for (int i = 1; i < n; ++i)
    if (col >= lat[i].offset) k = i;
  return k;
}

// -------------------------------------------------------------------------------------------
// rsample: z[s][d] from noise, FactorizedDistribution.rsample (mininf/nn.py:133-145)
// -------------------------------------------------------------------------------------------
// value of a constrained parameter from its unconstrained storage (ParameterizedDistribution.forward,
// mininf/nn.py:88-96: transform_to(real) is the identity, transform_to(positive) is exp)
__device__ __forceinline__ float apply_transform(uint8_t code, float raw) {
  return (code & MNF_T_MASK) == MNF_T_EXP ? expf(raw) : raw;
}

// `raw` != NULL (fused SVI step): the constrained parameters are computed here from the
// unconstrained ones, raw[d] -> p0 of column d, raw[D + d] -> p1, and written to `constrained`
// [2D] (where the latent table's p0 / p1 pointers point) for the later kernels of the step.
__global__ void rsample_kernel(const mnf_latent_t* __restrict__ lat, int n_lat, int S, int D,
                               const float* __restrict__ noise_in, uint64_t seed, uint64_t offset,
                               const uint64_t* __restrict__ offset_dev,
                               float* __restrict__ z, float* __restrict__ noise_out,
                               double* __restrict__ acc, uint32_t* __restrict__ status,
                               const float* __restrict__ raw, const uint8_t* __restrict__ tcode,
                               float* __restrict__ constrained) {
  if (offset_dev != nullptr) offset += *offset_dev;   // device-side call index (CUDA-graph replays)
  const int64_t n_z = (int64_t)S * D;
  const int64_t n_acc = (int64_t)S * (D + 1);
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nth = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n_acc; i += nth) acc[i] = 0.0;
  uint32_t bad = 0;
  for (int64_t i = tid; i < n_z; i += nth) {
    const int d = (int)(i % D);
    const mnf_latent_t L = lat[find_latent(lat, n_lat, d)];
    const int e = d - L.offset;
    float p0, p1;
    if (raw != nullptr) {
      p0 = apply_transform(tcode[d], raw[d]);
      p1 = apply_transform(tcode[D + d], raw[D + d]);
      if (i < D) {             // particle 0 publishes the constrained values
        constrained[d] = p0;
        constrained[D + d] = p1;
      }
    } else {
      p0 = L.p0[e];
      p1 = L.p1[e];
    }
    float nz, val;
    if (L.family == MNF_NORMAL) {
      // Normal.rsample: loc + eps * scale                       TORCH normal.py:82-85
      if (noise_in != nullptr) {
        nz = noise_in[i];
      } else {
        Philox rng(seed, offset, kPhiloxNormal | (uint64_t)i);
        const uint4 r = rng.next();
        nz = box_muller(r.x, r.y).x;
      }
      val = fmaf(nz, p1, p0);
      if (!(p1 > 0.0f) || p0 != p0) bad |= MNF_ST_BAD_PARAM;
    } else if (L.family == MNF_GAMMA) {
      // Gamma.rsample: standard_gamma(alpha) / rate, clamped at tiny   TORCH gamma.py:79-87
      nz = noise_in != nullptr ? noise_in[i] : philox_standard_gamma(p0, seed, offset, (uint64_t)i);
      val = fmaxf(nz / p1, kFloatTiny);
      if (!(p0 > 0.0f) || !(p1 > 0.0f)) bad |= MNF_ST_BAD_PARAM;
    } else {
      // Beta.rsample: first component of a 2-simplex Dirichlet draw    TORCH beta.py:84-85,
      // i.e. G1 / (G1 + G0) for independent standard gammas (dirichlet.py:85-88)
      if (noise_in != nullptr) {
        nz = noise_in[i];
      } else {
        const float g1 = philox_standard_gamma(p0, seed, offset, (uint64_t)(2 * i));
        const float g0 = philox_standard_gamma(p1, seed, offset, (uint64_t)(2 * i + 1) | (kPhiloxBeta0 >> 8));
        nz = fminf(fmaxf(g1 / (g1 + g0), kFloatEps), 1.0f - kFloatEps);
      }
      val = nz;
      if (!(p0 > 0.0f) || !(p1 > 0.0f)) bad |= MNF_ST_BAD_PARAM;
    }
    z[i] = val;
    noise_out[i] = nz;
  }
  if (bad) atomicOr(status, bad);
}

// -------------------------------------------------------------------------------------------
// small sites: LogProbTracer.sample + contribution (mininf/core.py:211-273) for short sites.
// grid = (blocks over elements, n_sites, particles): one block evaluates a chunk of one site for
// one particle; sums are reduced in the block and leave it as one fp64 atomic per target.
// -------------------------------------------------------------------------------------------
constexpr int kSmallThreads = 128;

// Sums of one thread over its elements of one (site, particle) pair; element-wise latent targets
// are added to acc directly (distinct addresses within a site; other sites may touch the same
// column, hence atomics).
struct SmallSums {
  double lp;
  float gA0, gB0, gA1, gB1;
};

__device__ inline SmallSums small_site_elements(const mnf_site_t& site, const float* zs, double* as, int64_t i0,
                                                int64_t stride, uint32_t& bad) {
  const bool g0 = link_has_latent(site.param[0]);
  const bool g1 = link_has_latent(site.param[1]);
  const bool two = site.family <= MNF_BETA;  // families with a second parameter
  const float w = (float)site.scale;
  SmallSums o;
  o.lp = 0.0; o.gA0 = 0.f; o.gB0 = 0.f; o.gA1 = 0.f; o.gB1 = 0.f;
  for (int64_t i = i0; i < site.numel; i += stride) {
    if (site.mask != nullptr && site.mask[i] == 0) continue;
    const float v = site.value_lat >= 0 ? zs[site.value_lat + i] : site.value[i];
    const LinkVal l0 = eval_link(site.param[0], zs, i);
    LinkVal l1; l1.value = 0.f; l1.du = 0.f; l1.x = 1.f;
    if (two) l1 = eval_link(site.param[1], zs, i);
    const Dens dn = density(site.family, v, l0.value, l1.value, g0 || g1);
    if (dn.bad_param) bad |= MNF_ST_BAD_PARAM;
    if (dn.bad_value) bad |= MNF_ST_BAD_VALUE;
    o.lp += (double)dn.lp;
    if (site.value_lat >= 0) atomicAdd(as + 1 + site.value_lat + i, (double)(w * dn.dv));
    if (g0) {
      const float du = dn.d0 * l0.du;
      const mnf_link_t& L = site.param[0];
      if (L.a_lat >= 0) { if (L.a_stride) atomicAdd(as + 1 + L.a_lat + L.a_stride * i, (double)(w * du)); else o.gA0 += du; }
      if (L.b_lat >= 0) { if (L.b_stride) atomicAdd(as + 1 + L.b_lat + L.b_stride * i, (double)(w * du * l0.x)); else o.gB0 += du * l0.x; }
    }
    if (two && g1) {
      const float du = dn.d1 * l1.du;
      const mnf_link_t& L = site.param[1];
      if (L.a_lat >= 0) { if (L.a_stride) atomicAdd(as + 1 + L.a_lat + L.a_stride * i, (double)(w * du)); else o.gA1 += du; }
      if (L.b_lat >= 0) { if (L.b_stride) atomicAdd(as + 1 + L.b_lat + L.b_stride * i, (double)(w * du * l1.x)); else o.gB1 += du * l1.x; }
    }
  }
  return o;
}

// acc column of sum `k` (0 log-density, 1..4 scalar-latent gradients) of a site, or -1
__device__ __forceinline__ int small_site_target(const mnf_site_t& site, int k) {
  const bool g0 = link_has_latent(site.param[0]);
  const bool g1 = link_has_latent(site.param[1]);
  const bool two = site.family <= MNF_BETA;
  const mnf_link_t& L0 = site.param[0];
  const mnf_link_t& L1 = site.param[1];
  if (k == 0) return 0;
  if (k == 1 && g0 && L0.a_lat >= 0 && L0.a_stride == 0) return 1 + L0.a_lat;
  if (k == 2 && g0 && L0.b_lat >= 0 && L0.b_stride == 0) return 1 + L0.b_lat;
  if (k == 3 && two && g1 && L1.a_lat >= 0 && L1.a_stride == 0) return 1 + L1.a_lat;
  if (k == 4 && two && g1 && L1.b_lat >= 0 && L1.b_stride == 0) return 1 + L1.b_lat;
  return -1;
}

__global__ void __launch_bounds__(kSmallThreads)
small_sites_kernel(const mnf_site_t* __restrict__ sites, const float* __restrict__ z, int S, int D,
                   double* __restrict__ acc, uint32_t* __restrict__ status) {
  const mnf_site_t site = sites[blockIdx.y];
  const int s = blockIdx.z;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if ((int64_t)blockIdx.x * blockDim.x >= site.numel) return;
  const float* zs = z + (int64_t)s * D;
  double* as = acc + (int64_t)s * (D + 1);
  uint32_t bad = 0;
  SmallSums t = small_site_elements(site, zs, as, i0, stride, bad);
  // block reduction: warp shuffles, then one warp finishes through shared memory
  __shared__ double s_red[kSmallThreads / 32][5];
  t.lp = warp_sum(t.lp);
  t.gA0 = warp_sum(t.gA0); t.gB0 = warp_sum(t.gB0); t.gA1 = warp_sum(t.gA1); t.gB1 = warp_sum(t.gB1);
  if ((threadIdx.x & 31) == 0) {
    double* row = s_red[threadIdx.x >> 5];
    row[0] = t.lp; row[1] = t.gA0; row[2] = t.gB0; row[3] = t.gA1; row[4] = t.gB1;
  }
  __syncthreads();
  if (threadIdx.x < 5) {
    double total = 0.0;
    for (int k = 0; k < kSmallThreads / 32; ++k) total += s_red[k][threadIdx.x];
    const int target = small_site_target(site, threadIdx.x);
    if (target >= 0 && total != 0.0) atomicAdd(as + target, site.scale * total);
  }
  if (bad) atomicOr(status, bad);
}

// The same evaluation inside a single block (the fused tail kernel below): one warp per
// (site, particle) pair, lanes over the elements.
__device__ inline void small_sites_block(const mnf_site_t* __restrict__ sites, int n_sites, const float* __restrict__ z,
                                         int S, int D, double* __restrict__ acc, uint32_t* __restrict__ status) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  uint32_t bad = 0;
  // Short sites (the scalar priors of most models): one THREAD per (site, particle), all pairs at once.
  // A warp per pair left 31 lanes idle and walked the pairs in rounds of n_warps, each round a chain of
  // dependent loads and five fp64 shuffle reductions: 5 scalar priors x 64 particles took 38 us of the
  // C5 tail kernel (75 000 cycles); this pass takes one round.
  constexpr int64_t kShortSite = 4;
  for (int task = threadIdx.x; task < n_sites * S; task += blockDim.x) {
    const mnf_site_t site = sites[task / S];
    if (site.numel > kShortSite) continue;
    const int s = task % S;
    double* as = acc + (int64_t)s * (D + 1);
    SmallSums t = small_site_elements(site, z + (int64_t)s * D, as, 0, 1, bad);
    const double sums[5] = {t.lp, (double)t.gA0, (double)t.gB0, (double)t.gA1, (double)t.gB1};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const int target = small_site_target(site, k);
      if (target >= 0 && sums[k] != 0.0) atomicAdd(as + target, site.scale * sums[k]);
    }
  }
  for (int task = warp; task < n_sites * S; task += n_warps) {
    const mnf_site_t site = sites[task / S];
    if (site.numel <= kShortSite) continue;
    const int s = task % S;
    const float* zs = z + (int64_t)s * D;
    double* as = acc + (int64_t)s * (D + 1);
    SmallSums t = small_site_elements(site, zs, as, lane, 32, bad);
    double sums[5] = {t.lp, (double)t.gA0, (double)t.gB0, (double)t.gA1, (double)t.gB1};
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      const double total = warp_sum(sums[k]);
      const int target = small_site_target(site, k);
      if (lane == 0 && target >= 0 && total != 0.0) atomicAdd(as + target, site.scale * total);
    }
  }
  if (bad) atomicOr(status, bad);
}

// -------------------------------------------------------------------------------------------
// Fixed-order reduction of per-CTA sweep partials [n_cta][S][ncol] (fp32) into acc (fp64).
// Column c of a partial goes to acc column colmap(c): 0 -> log-density, 1+k -> latent column.
// -------------------------------------------------------------------------------------------
constexpr int kReduceThreads = 256;

// One warp per (particle, column) output: lane l sums CTAs l, l+32, ... in order, then a fixed
// butterfly combines the lanes - the same order on every run, and no serial walk over the CTAs.
__global__ void __launch_bounds__(kReduceThreads)
reduce_partials_kernel(const float* __restrict__ partial, int n_cta, int S, int ncol,
                       ColMap map, double weight, int D, double* __restrict__ acc) {
  const int idx = blockIdx.x * (kReduceThreads / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (idx >= S * ncol) return;
  const int s = idx / ncol, c = idx % ncol;
  int target;
  if (c == 0) target = 0;
  else if (c <= map.n_vec) target = 1 + map.vec_lat + (c - 1);
  else {
    const int k = c - 1 - map.n_vec;
    if (k >= map.n_scalar || map.scalar_lat[k] < 0) return;
    target = 1 + map.scalar_lat[k];
  }
  double t = 0.0;
  for (int b = lane; b < n_cta; b += 32) t += (double)partial[((int64_t)b * S + s) * ncol + c];
  t = warp_sum(t);
  // several partial columns may map to one latent column (e.g. the same scalar latent used by
  // two fused sites), hence the atomic; launches are stream-ordered.
  if (lane == 0) atomicAdd(acc + (int64_t)s * (D + 1) + target, weight * t);
}

// -------------------------------------------------------------------------------------------
// finalize: loss = -(mean_s acc[s][0] + H[q]) (mininf/nn.py:226-228) and gradients w.r.t. the
// constrained parameters via the pathwise derivative of each family's rsample.
// -------------------------------------------------------------------------------------------
constexpr int kFinalThreads = 512;

// torch.optim.Adam (defaults: no weight decay, no amsgrad) over the unconstrained parameters of
// the packed latent sites, fused into the last kernel of the step (README.md:63-69's
// `optimizer.step()`; ParameterizedDistribution's transforms, mininf/nn.py:88-96, are chained
// here instead of by autograd). Passed by value; `raw == nullptr` means "no optimiser".
struct AdamArgs {
  float lr, beta1, beta2, eps;
  float* raw;                 // [2D]
  const uint8_t* transform;   // [2D] MNF_T_* (| MNF_T_FROZEN)
  float* m;                   // [2D]
  float* v;                   // [2D]
  int64_t* step;              // device: updates applied so far
};

__device__ __forceinline__ void adam_update(const AdamArgs& ad, int idx, float constrained, double grad_constrained,
                                            double bc1, double bc2_sqrt) {
  const uint8_t code = ad.transform[idx];
  if (code & MNF_T_FROZEN) return;
  // chain rule through the transform: d exp(raw)/d raw = the constrained value itself
  const float g = (float)((code & MNF_T_MASK) == MNF_T_EXP ? grad_constrained * (double)constrained : grad_constrained);
  const float m = ad.beta1 * ad.m[idx] + (1.0f - ad.beta1) * g;
  const float v = ad.beta2 * ad.v[idx] + (1.0f - ad.beta2) * g * g;
  ad.m[idx] = m;
  ad.v[idx] = v;
  const float denom = (float)((double)sqrtf(v) / bc2_sqrt) + ad.eps;
  ad.raw[idx] -= (float)((double)ad.lr / bc1) * (m / denom);
}

// One warp per latent column: the lanes walk the particles (the implicit Gamma / Beta gradients
// are ~100 fp64 operations per particle), a fixed butterfly combines them, lane 0 adds the
// entropy terms. Sums run in the same order on every launch. Works for any block size that is a
// multiple of 32 (<= 1024); acc is read past L1 (the fused tail kernel updates it with atomics).
template <bool STAGED>
__device__ __forceinline__ double load_acc(const double* p) {
  if constexpr (STAGED) return *p;          // staged in shared memory by the tail kernel
  else return __ldcg(p);
}

// Gamma / Beta columns of finalize_block, kept OUT OF LINE: the tail kernel runs once per step with a
// cold instruction cache, and its time follows its code size (22 K instructions: 32 us at C2; with the
// fp64 special functions and this block out of line: 13 K and 24 us). One WARP per column, lanes over
// the particles (the implicit reparameterisation gradients are ~100 fp64 operations per particle).
template <bool STAGED>
static __device__ __noinline__ void finalize_other_columns(const mnf_latent_t* __restrict__ lat, int n_lat, int S, int D,
                                                           const float* noise, const double* acc, int with_entropy,
                                                           float* __restrict__ out, const AdamArgs& adam, double bc1,
                                                           double bc2_sqrt, double* ent_io, bool* nonfinite_io) {
  // One column at a time, the whole block on it: the particles spread over the threads (one fp64
  // implicit-gradient evaluation per thread instead of S / 32 in a row per lane: the three regimes of
  // standard_gamma_grad diverge inside a warp, so a warp pays for all of them per round), the entropy
  // terms (lgamma / digamma / trigamma) on the last warp at the same time. Fixed-order sums.
  __shared__ double s_part[32][2];
  __shared__ double s_h[3];
  const int kWarps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double invS = 1.0 / (double)S;
  double ent = 0.0;
  bool nonfinite = false;
  for (int d = 0; d < D; ++d) {
    const mnf_latent_t L = lat[find_latent(lat, n_lat, d)];
    if (L.family == MNF_NORMAL) continue;          // uniform over the block
    const int e = d - L.offset;
    const double p0 = (double)L.p0[e], p1 = (double)L.p1[e];
    double g0 = 0.0, g1 = 0.0;
    if (L.family == MNF_GAMMA) {
      // z = g/rate: dz/dalpha = standard_gamma_grad(alpha, g)/rate, dz/drate = -g/rate^2
      // (the clamp_ at tiny is outside autograd: gradients as if unclamped)
      for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const double g = load_acc<STAGED>(acc + (int64_t)s * (D + 1) + 1 + d);
        const double gam = (double)noise[(int64_t)s * D + d];
        g0 += g * standard_gamma_grad(p0, gam) / p1;
        g1 += g * (-gam / (p1 * p1));
      }
    } else {
      // Beta(c1 = p0, c0 = p1) as a 2-simplex Dirichlet; _Dirichlet backward dirichlet.py:16-35:
      // grad_k = dirichlet_grad(x_k, c_k, total) * (go_k - sum_j x_j go_j) with go = (g, 0).
      const double tot = p0 + p1;
      for (int s = threadIdx.x; s < S; s += blockDim.x) {
        const double g = load_acc<STAGED>(acc + (int64_t)s * (D + 1) + 1 + d);
        const double x = (double)noise[(int64_t)s * D + d];
        g0 += dirichlet_grad(x, p0, tot) * g * (1.0 - x);
        g1 += dirichlet_grad(1.0 - x, p1, tot) * (-x * g);
      }
    }
    g0 = warp_sum(g0);
    g1 = warp_sum(g1);
    if (lane == 0) { s_part[warp][0] = g0; s_part[warp][1] = g1; }
    if (warp == kWarps - 1 && lane == 0) {
      double h = 0.0, dh0 = 0.0, dh1 = 0.0;
      if (L.family == MNF_GAMMA) {
        // H = alpha - log(rate) + lgamma(alpha) + (1-alpha) digamma(alpha)   TORCH gamma.py:100-106
        h = p0 - log(p1) + lgamma(p0) + (1.0 - p0) * digamma_d(p0);
        dh0 = 1.0 + (1.0 - p0) * trigamma_d(p0);
        dh1 = -1.0 / p1;
      } else {
        // Dirichlet entropy with k = 2                                TORCH dirichlet.py:122-130
        const double tot = p0 + p1;
        const double dt = digamma_d(tot);
        h = lgamma(p0) + lgamma(p1) - lgamma(tot) - (p0 - 1.0) * digamma_d(p0) -
            (p1 - 1.0) * digamma_d(p1) + (tot - 2.0) * dt;
        const double tt = trigamma_d(tot);
        dh0 = -(p0 - 1.0) * trigamma_d(p0) + (tot - 2.0) * tt;
        dh1 = -(p1 - 1.0) * trigamma_d(p1) + (tot - 2.0) * tt;
      }
      if (!with_entropy) { h = 0.0; dh0 = 0.0; dh1 = 0.0; }
      s_h[0] = h; s_h[1] = dh0; s_h[2] = dh1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      g0 = 0.0; g1 = 0.0;
      for (int w = 0; w < kWarps; ++w) { g0 += s_part[w][0]; g1 += s_part[w][1]; }
      ent += s_h[0];
      const double o0 = -(g0 * invS + s_h[1]);
      const double o1 = -(g1 * invS + s_h[2]);
      if (!isfinite(o0) || !isfinite(o1)) nonfinite = true;
      out[1 + d] = (float)o0;
      out[1 + D + d] = (float)o1;
      if (adam.raw != nullptr) {
        adam_update(adam, d, (float)p0, o0, bc1, bc2_sqrt);
        adam_update(adam, D + d, (float)p1, o1, bc1, bc2_sqrt);
      }
    }
    __syncthreads();     // s_part / s_h are reused by the next column
  }
  *ent_io += ent;
  *nonfinite_io = *nonfinite_io || nonfinite;
}

template <bool STAGED>
__device__ inline void finalize_block(const mnf_latent_t* __restrict__ lat, int n_lat, int S, int D,
                                      const float* z, const float* noise,
                                      const double* acc, int with_entropy, float* __restrict__ out,
                                      uint64_t* __restrict__ step_counter, uint32_t* __restrict__ status,
                                      const AdamArgs& adam) {
  __shared__ double red[32];
  const int kWarps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // last kernel of a step: the next replay of a captured step draws with the next call index
  if (step_counter != nullptr && threadIdx.x == 0) *step_counter += 1;
  // Adam's bias corrections of this step: one thread computes them (fp64 pow), everyone reads them
  __shared__ double bias[2];
  if (threadIdx.x == 0) {
    bias[0] = 1.0;
    bias[1] = 1.0;
    if (adam.raw != nullptr) {
      // 1 - beta^t = -expm1(t log beta): expm1f keeps the small-t digits (relative error ~2e-7, far
      // below what a step of lr * m / sqrt(v) resolves); the fp64 versions were ~3000 dependent cycles
      // in front of a block-wide barrier
      const float t = (float)(*adam.step + 1);
      bias[0] = (double)-expm1f(t * logf(adam.beta1));
      bias[1] = (double)sqrtf(-expm1f(t * logf(adam.beta2)));
    }
  }
  __syncthreads();
#ifdef MNF_TAIL_DEBUG
  const long long f0 = clock64();
#endif
  const double bc1 = bias[0], bc2_sqrt = bias[1];
  const double invS = 1.0 / (double)S;
  double ent = 0.0;        // this thread's share of the entropy (+ warp 0: the log joint)
  bool nonfinite = false;
  // ---- Normal columns: one THREAD per column. The particle loop reads acc[s][1+d] and noise[s][d]
  // coalesced across the threads of a warp and is unrolled, so the whole block pays a handful of
  // L2 round trips instead of one dependent chain per column (the step is latency-bound here).
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const mnf_latent_t L = lat[find_latent(lat, n_lat, d)];
    if (L.family != MNF_NORMAL) continue;
    const int e = d - L.offset;
    const double p0 = (double)L.p0[e], p1 = (double)L.p1[e];
    double g0 = 0.0, g1 = 0.0;
    int s = 0;
    for (; s + 8 <= S; s += 8) {
      double g[8];
      float nz[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        g[u] = load_acc<STAGED>(acc + (int64_t)(s + u) * (D + 1) + 1 + d);
        nz[u] = noise[(int64_t)(s + u) * D + d];
      }
      // z = loc + eps*scale; pairwise sums: short dependency chains (fixed order, deterministic)
      const double a01 = g[0] + g[1], a23 = g[2] + g[3], a45 = g[4] + g[5], a67 = g[6] + g[7];
      const double b01 = g[0] * (double)nz[0] + g[1] * (double)nz[1], b23 = g[2] * (double)nz[2] + g[3] * (double)nz[3];
      const double b45 = g[4] * (double)nz[4] + g[5] * (double)nz[5], b67 = g[6] * (double)nz[6] + g[7] * (double)nz[7];
      g0 += (a01 + a23) + (a45 + a67);
      g1 += (b01 + b23) + (b45 + b67);
    }
    for (; s < S; ++s) {
      const double g = load_acc<STAGED>(acc + (int64_t)s * (D + 1) + 1 + d);
      g0 += g;
      g1 += g * (double)noise[(int64_t)s * D + d];
    }
    // H = 0.5 + 0.5 log(2 pi) + log(scale)                      TORCH normal.py:114-115
    double h = 0.5 + 0.91893853320467274178 + (double)logf((float)p1), dh1 = 1.0 / p1;
    if (!with_entropy) { h = 0.0; dh1 = 0.0; }
    ent += h;
    const double o0 = -(g0 * invS);
    const double o1 = -(g1 * invS + dh1);
    if (!isfinite(o0) || !isfinite(o1)) nonfinite = true;
    out[1 + d] = (float)o0;
    out[1 + D + d] = (float)o1;
    if (adam.raw != nullptr) {
      adam_update(adam, d, (float)p0, o0, bc1, bc2_sqrt);
      adam_update(adam, D + d, (float)p1, o1, bc1, bc2_sqrt);
    }
  }
#ifdef MNF_TAIL_DEBUG
  __syncthreads();
  const long long f1 = clock64();
#endif
  // ---- Gamma / Beta columns (out of line, skipped when every latent is Normal)
  {
    bool any_other = false;
    for (int i = 0; i < n_lat; ++i) any_other = any_other || lat[i].family != MNF_NORMAL;
    if (any_other)
      finalize_other_columns<STAGED>(lat, n_lat, S, D, noise, acc, with_entropy, out, adam, bc1, bc2_sqrt, &ent, &nonfinite);
  }
#ifdef MNF_TAIL_DEBUG
  __syncthreads();
  const long long f2 = clock64();
  if (threadIdx.x == 0) printf("finalize: normal columns %lld  other columns %lld cycles\n", f1 - f0, f2 - f1);
#endif
  // total log joint over particles: warp 0's lanes, added to its entropy share
  if (warp == 0) {
    double lj = 0.0;
    for (int s = lane; s < S; s += 32) lj += load_acc<STAGED>(acc + (int64_t)s * (D + 1));
    lj = warp_sum(lj);
    if (lane == 0) ent += lj * invS;
  }
  ent = warp_sum(ent);      // every lane may hold a share now (fixed butterfly: same order every launch)
  if (lane == 0) red[warp] = ent;
  __syncthreads();
  if (threadIdx.x == 0) {
    double total = 0.0;
    for (int w = 0; w < kWarps; ++w) total += red[w];
    const double loss = -total;
    out[0] = (float)loss;
    if (!isfinite(loss)) nonfinite = true;
    if (adam.raw != nullptr) *adam.step += 1;
  }
  if (nonfinite) atomicOr(status, MNF_ST_NONFINITE);
}

__global__ void __launch_bounds__(kFinalThreads)
finalize_kernel(const mnf_latent_t* __restrict__ lat, int n_lat, int S, int D,
                const float* __restrict__ z, const float* __restrict__ noise,
                const double* __restrict__ acc, int with_entropy, float* __restrict__ out,
                uint64_t* __restrict__ step_counter, uint32_t* __restrict__ status) {
  AdamArgs none;
  none.raw = nullptr;
  finalize_block<false>(lat, n_lat, S, D, z, noise, acc, with_entropy, out, step_counter, status, none);
}

// -------------------------------------------------------------------------------------------
// Cross-rank exchange of the step accumulator over peer memory (NVLink / NVSwitch), replacing the
// ncclAllReduce of SURVEY §8e: every rank PUSHES its partial accumulator into a per-source inbox
// on each peer and raises that peer's arrival flag; the tail kernel of each rank then waits for
// its W - 1 flags and adds the inboxes in rank order - the same order on every rank, so all
// ranks hold bit-identical totals and parameters stay replicated without a broadcast. Inboxes
// are double-buffered by the parity of the exchange epoch: a rank can write into a peer's slot
// of epoch k + 2 only after that peer signalled epoch k + 1, i.e. after it finished reading
// epoch k. No host call, no NCCL launch: the step stays capturable in a CUDA graph.
// -------------------------------------------------------------------------------------------
struct XrankArgs {
  int world, rank;
  int64_t n;                    // doubles per accumulator
  double* inbox;                // local [2][world][n]
  uint64_t* flags;              // local [world], flags[r] = last epoch rank r delivered
  double* const* peer_inbox;    // device array [world]: inbox base on every rank
  uint64_t* const* peer_flags;  // device array [world]
  uint64_t* epoch;              // local device word: exchanges completed so far
};

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

constexpr int kXrankThreads = 512;

// grid = world: block b delivers this rank's accumulator to rank b
__global__ void __launch_bounds__(kXrankThreads)
xrank_push_kernel(XrankArgs xr, const double* __restrict__ acc) {
  const int peer = blockIdx.x;
  if (peer == xr.rank) return;
  const uint64_t epoch = *xr.epoch + 1;
  double* dst = xr.peer_inbox[peer] + ((epoch & 1) * (uint64_t)xr.world + (uint64_t)xr.rank) * (uint64_t)xr.n;
  for (int64_t i = threadIdx.x; i < xr.n; i += kXrankThreads) dst[i] = acc[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(xr.peer_flags[peer] + xr.rank, epoch);
}

// wait for every peer's delivery of this epoch, then acc += inboxes in rank order (one block)
// `dst`: where the totals go (acc itself, or the tail kernel's shared-memory copy)
__device__ inline void xrank_gather_block(const XrankArgs& xr, const double* __restrict__ acc, double* dst,
                                          uint32_t* __restrict__ status) {
  const uint64_t epoch = *xr.epoch + 1;
  if ((int)threadIdx.x < xr.world && (int)threadIdx.x != xr.rank) {
    // bounded wait (about two seconds): a dead peer must not hang the GPU
    const long long start = clock64();
    while (ld_acquire_sys(xr.flags + threadIdx.x) < epoch) {
      if (clock64() - start > 4000000000LL) {
        atomicOr(status, MNF_ST_XRANK_TIMEOUT);
        break;
      }
    }
  }
  __syncthreads();
  const double* slot = xr.inbox + (epoch & 1) * (uint64_t)xr.world * (uint64_t)xr.n;
  for (int64_t i = threadIdx.x; i < xr.n; i += blockDim.x) {
    double total = 0.0;
    for (int r = 0; r < xr.world; ++r)
      total += r == xr.rank ? acc[i] : __ldcv(slot + (uint64_t)r * (uint64_t)xr.n + i);
    dst[i] = total;
  }
  __syncthreads();
  if (threadIdx.x == 0) *xr.epoch = epoch;
}

// -------------------------------------------------------------------------------------------
// Tail of a step in ONE launch: [cross-rank gather] -> [latent-valued (prior) sites] -> finalize
// [-> Adam]. A single block: the work is O(S * D).
// -------------------------------------------------------------------------------------------
constexpr int kTailThreads = 512;    // 128 registers per thread for the fp64 implicit-gradient code

// bytes of dynamic shared memory the staged variant needs: acc [S][1+D] fp64, noise and z [S][D] fp32
inline size_t tail_stage_bytes(int S, int D) {
  return (size_t)S * (D + 1) * sizeof(double) + (size_t)2 * S * D * sizeof(float);
}

// STAGED: the step's O(S*D) state is first copied into shared memory by all threads at once (one
// round of coalesced L2 reads); every later phase - the prior sites' atomics included - then works
// on-chip. The tail is pure latency (a few hundred dependent L2 round trips otherwise), and at 8 GPUs
// under strong scaling every 10 us of it is 1.5 % of the step.
template <bool STAGED>
__global__ void __launch_bounds__(kTailThreads)
tail_kernel(XrankArgs xr, const mnf_site_t* __restrict__ global_sites, int n_global,
            const mnf_latent_t* __restrict__ lat, int n_lat, int S, int D, const float* __restrict__ z,
            const float* __restrict__ noise, double* __restrict__ acc, int with_entropy, float* __restrict__ out,
            uint64_t* __restrict__ step_counter, uint32_t* __restrict__ status, AdamArgs adam) {
  extern __shared__ __align__(16) unsigned char tail_smem[];
  // the latent table is read with dependent loads by every phase below (find_latent walks it): a copy
  // in shared memory takes those round trips to L2 off the kernel's latency chain
  constexpr int kLatStaged = 32;
  __shared__ mnf_latent_t s_lat[kLatStaged];
  if (n_lat <= kLatStaged) {
    for (int i = threadIdx.x; i < n_lat; i += kTailThreads) s_lat[i] = lat[i];
    lat = s_lat;      // visible after the first barrier below (every path has one before the table is read)
  }
  if constexpr (STAGED) {
    double* acc_s = reinterpret_cast<double*>(tail_smem);
    float* noise_s = reinterpret_cast<float*>(acc_s + (size_t)S * (D + 1));
    float* z_s = noise_s + (size_t)S * D;
#ifdef MNF_TAIL_DEBUG
    const long long t0 = clock64();
#endif
    if (xr.world > 1) {
      xrank_gather_block(xr, acc, acc_s, status);
    } else {
      for (int i = threadIdx.x; i < S * (D + 1); i += kTailThreads) acc_s[i] = __ldcg(acc + i);
    }
    for (int i = threadIdx.x; i < S * D; i += kTailThreads) {
      noise_s[i] = noise[i];
      z_s[i] = z[i];
    }
    __syncthreads();
#ifdef MNF_TAIL_DEBUG
    const long long t1 = clock64();
#endif
    if (n_global > 0) {
      small_sites_block(global_sites, n_global, z_s, S, D, acc_s, status);
      __syncthreads();
    }
#ifdef MNF_TAIL_DEBUG
    const long long t2 = clock64();
#endif
    finalize_block<true>(lat, n_lat, S, D, z_s, noise_s, acc_s, with_entropy, out, step_counter, status, adam);
#ifdef MNF_TAIL_DEBUG
    __syncthreads();
    if (threadIdx.x == 0) printf("tail: stage %lld  prior sites %lld  finalize %lld cycles (n_global %d, n_lat %d, S %d, D %d)\n", t1 - t0, t2 - t1, clock64() - t2, n_global, n_lat, S, D);
#endif
  } else {
    if (xr.world > 1) xrank_gather_block(xr, acc, acc, status);
    if (n_global > 0) {
      small_sites_block(global_sites, n_global, z, S, D, acc, status);
      __threadfence();
      __syncthreads();
    }
    finalize_block<false>(lat, n_lat, S, D, z, noise, acc, with_entropy, out, step_counter, status, adam);
  }
}

// -------------------------------------------------------------------------------------------
// integer-exact counting scan (parity checks on masks and count data)
// -------------------------------------------------------------------------------------------
__global__ void masked_count_kernel(const float* __restrict__ value, const uint8_t* __restrict__ mask,
                                    int64_t n, unsigned long long* __restrict__ out) {
  unsigned long long cnt = 0, tot = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const bool m = mask == nullptr || mask[i] != 0;
    if (m) {
      cnt += 1;
      tot += (unsigned long long)(long long)llrintf(value[i]);
    }
  }
  cnt = warp_sum(cnt);
  tot = warp_sum(tot);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out, cnt);
    atomicAdd(out + 1, tot);
  }
}

}  // namespace mnf
