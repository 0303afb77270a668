// Row-latent sweep: a per-observation latent matrix Z [n][p] (p <= 32) with a mean-field Normal
// approximation, and the prior / feature / response sites that touch it, for all particles in one
// pass (config C4: examples/regression-with-feature-uncertainty.md:28-38 widened to p features).
//
// Replaces, per step: Normal.rsample of n*p values per particle (TORCH normal.py:82-85), the
// element-wise log_prob chains of three sites (mininf/core.py:241), `z @ slope` and its backward,
// the entropy of q(Z) (TORCH normal.py:114-115) and autograd's n*p-sized gradient passes.
//
// Mapping: a warp owns a row, lane j owns feature j. Per row a lane keeps loc, scale, feature and
// its two gradient accumulators in registers and loops over the particles in groups of four (one
// Philox4x32-10 call -> two Box-Muller pairs). The per-particle dot products Z_i . beta_s are
// reduced AND transposed in one 31-shuffle butterfly so that lane s ends up with particle s's
// linear predictor: each lane then evaluates one exp per row instead of 32 redundant ones.
// Per-(particle, lane) statistics live in registers (SP is a template parameter), so the hot loop
// has no shared-memory or atomic traffic. Algorithmic bytes per row: p * (loc 4 + scale 4 +
// feature 4 + two gradients 8) + response 4.
#pragma once

#include "common.cuh"

namespace mnf {

constexpr int kRowThreads = 256;
constexpr int kRowWarps = kRowThreads / 32;

struct RowParticle {   // per-particle scalars of the row-latent sites, staged in shared memory
  float prior_loc, prior_inv_var, prior_scale, prior_dscale;  // Normal prior of Z
  float feat_inv_var, feat_scale, feat_dscale;                // Normal features
  float icpt, resp_scale, resp_dscale;
};

// partial layout per CTA: [S][ncol], ncol = 1 + p + 5:
//   0 log-density (+ entropy share), 1..p beta gradient, p+1 intercept, p+2 prior loc (du),
//   p+3 prior scale (du), p+4 feature scale (du), p+5 response scale (du)
template <int SP>
__global__ void __launch_bounds__(kRowThreads)
rowlatent_kernel(mnf_rowlatent_t d, const float* __restrict__ z, int S, int D, int s_begin,
                 int first_pass, uint64_t seed, uint64_t offset, const uint64_t* __restrict__ offset_dev,
                 int with_entropy, float* __restrict__ partial, uint32_t* __restrict__ status) {
  extern __shared__ float smem[];
  if (offset_dev != nullptr) offset += *offset_dev;   // device-side call index (CUDA-graph replays)
  const int p = d.p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sBeta = smem;                                             // [SP][32]
  RowParticle* sPar = reinterpret_cast<RowParticle*>(sBeta + SP * 32);   // [SP]
  float* sOut = reinterpret_cast<float*>(sPar + SP);               // [kRowWarps][SP][32 + 8] reduction scratch

  const int s_count = min(SP, S - s_begin);   // particles handled by this launch
  for (int i = threadIdx.x; i < SP * 32; i += kRowThreads) {
    const int s = i >> 5, j = i & 31;
    sBeta[i] = (s < s_count && j < p && d.resp != nullptr) ? z[(int64_t)(s_begin + s) * D + d.beta_lat + j] : 0.0f;
  }
  uint32_t bad = 0;
  for (int s = threadIdx.x; s < SP; s += kRowThreads) {
    RowParticle rp;
    rp.prior_loc = 0.f; rp.prior_inv_var = 0.f; rp.prior_scale = 1.f; rp.prior_dscale = 0.f;
    rp.feat_inv_var = 0.f; rp.feat_scale = 1.f; rp.feat_dscale = 0.f;
    rp.icpt = 0.f; rp.resp_scale = 1.f; rp.resp_dscale = 0.f;
    if (s < s_count) {
      const float* zs = z + (int64_t)(s_begin + s) * D;
      const LinkVal pl = eval_link(d.prior_loc, zs, 0), ps = eval_link(d.prior_scale, zs, 0);
      rp.prior_loc = pl.value; rp.prior_scale = ps.value; rp.prior_dscale = ps.du;
      rp.prior_inv_var = 1.0f / (ps.value * ps.value);
      if (!(ps.value > 0.0f)) bad |= MNF_ST_BAD_PARAM;
      if (d.feat != nullptr) {
        const LinkVal fs = eval_link(d.feat_scale, zs, 0);
        rp.feat_scale = fs.value; rp.feat_dscale = fs.du; rp.feat_inv_var = 1.0f / (fs.value * fs.value);
        if (!(fs.value > 0.0f)) bad |= MNF_ST_BAD_PARAM;
      }
      if (d.resp != nullptr) {
        rp.icpt = d.icpt_const + (d.icpt_lat >= 0 ? zs[d.icpt_lat] : 0.0f);
        if (d.resp_family == MNF_NORMAL) {
          const LinkVal rs = eval_link(d.resp_scale, zs, 0);
          rp.resp_scale = rs.value; rp.resp_dscale = rs.du;
          if (!(rs.value > 0.0f)) bad |= MNF_ST_BAD_PARAM;
        }
      }
    }
    sPar[s] = rp;
  }
  __syncthreads();

  // per-(particle, lane) statistics, registers
  float z2[SP], gb[SP];      // sum (z - prior_loc)^2 ; sum dlp/deta * z
#pragma unroll
  for (int s = 0; s < SP; ++s) { z2[s] = 0.f; gb[s] = 0.f; }
  // per-particle statistics owned by lane s (after the butterfly lane s holds particle s)
  float ga = 0.f, lpy = 0.f, gresp_scale = 0.f;
  // The feature scale and the prior location are constants in this build (checked on the host),
  // so their residual sums need no per-particle breakdown: only the mean over particles is used.
  float f2 = 0.f;            // sum over particles and rows of (x - z)^2
  double ent = 0.0;          // entropy of q(Z)
  float n_rows_lane = 0.f;
  bool bad_value = false;
  const float invS = 1.0f / (float)S;
  const bool active = lane < p;

  const int64_t warp_global = (int64_t)blockIdx.x * kRowWarps + warp;
  const int64_t warps_total = (int64_t)gridDim.x * kRowWarps;
  for (int64_t row = warp_global; row < d.n_rows; row += warps_total) {
    const int64_t e = row * p + lane;
    const float loc = active ? __ldg(d.loc + e) : 0.f;
    const float scale = active ? __ldg(d.scale + e) : 1.f;
    const float x = (active && d.feat != nullptr) ? __ldg(d.feat + e) : 0.f;
    const float y = d.resp != nullptr ? __ldg(d.resp + row) : 0.f;
    if (active && !(scale > 0.0f)) bad |= MNF_ST_BAD_PARAM;
    if (y != y || x != x) bad_value = true;
    float gl = 0.f, gs = 0.f;
    float zs[SP], v[SP];
    n_rows_lane += 1.f;
    // ---- draws and the element-wise sites -------------------------------------------------
#pragma unroll
    for (int q = 0; q < SP / 4; ++q) {
      float eps4[4];
      if (d.eps != nullptr) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int s = s_begin + 4 * q + t;
          eps4[t] = (active && s < S) ? __ldg(d.eps + ((int64_t)s * d.n_rows + row) * p + lane) : 0.f;
        }
      } else {
        Philox rng(seed, offset, ((uint64_t)e << 8) | (uint64_t)((s_begin >> 2) + q));
        const uint4 r = rng.next();
        const float2 n0 = box_muller(r.x, r.y), n1 = box_muller(r.z, r.w);
        eps4[0] = n0.x; eps4[1] = n0.y; eps4[2] = n1.x; eps4[3] = n1.y;
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int s = 4 * q + t;
        const RowParticle& rp = sPar[s];
        const float eps = eps4[t];
        const float zz = active ? fmaf(eps, scale, loc) : 0.f;
        zs[s] = zz;
        const float live = (active && s < s_count) ? 1.f : 0.f;
        // prior: d/dz = -(z - m)/ps^2
        const float dzp = zz - rp.prior_loc;
        z2[s] = fmaf(live * dzp, dzp, z2[s]);
        float dz = -dzp * rp.prior_inv_var;
        // features: d/dz = (x - z)/ns^2
        if (d.feat != nullptr) {
          const float r = x - zz;
          f2 = fmaf(live * r, r, f2);
          dz = fmaf(r, rp.feat_inv_var, dz);
        }
        dz *= live;
        gl += dz;
        gs = fmaf(dz, eps, gs);
        v[s] = zz * sBeta[s * 32 + lane];
      }
    }
    // ---- response: transpose-reduce the dot products so lane s owns particle s -----------------
    if (d.resp != nullptr) {
#pragma unroll
      for (int w = SP / 2; w >= 1; w >>= 1) {
        // SP < 32 leaves the upper lanes as idle copies; the butterfly still lands particle s on lane s
        const bool upper = (lane & w) != 0;
#pragma unroll
        for (int k = 0; k < w; ++k) {
          const float send = upper ? v[k] : v[k + w];
          const float keep = upper ? v[k + w] : v[k];
          v[k] = keep + __shfl_xor_sync(0xffffffffu, send, w);
        }
      }
      // lanes beyond SP hold partial sums of other lane groups: fold them in
#pragma unroll
      for (int w = SP; w < 32; w <<= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], w);
      const int sl = lane & (SP - 1);
      const RowParticle& rp = sPar[sl];
      const float eta = rp.icpt + v[0];
      float deta = 0.f, lp = 0.f, dsc = 0.f;
      if (sl < s_count) {
        if (d.resp_family == MNF_POISSON) {
          if (d.resp_transform == MNF_T_EXP) {
            const float rate = expf(eta);
            lp = fmaf(y, eta, -rate) - lgammaf(y + 1.0f);
            deta = y - rate;
          } else {
            lp = xlogy(y, eta) - eta - lgammaf(y + 1.0f);
            deta = y == 0.f ? -1.f : y / eta - 1.0f;
          }
        } else if (d.resp_family == MNF_NORMAL) {
          const float inv = 1.0f / rp.resp_scale;
          const float r = (y - eta) * inv;
          lp = -0.5f * r * r - logf(rp.resp_scale) - kLogSqrt2Pi;
          deta = r * inv;
          dsc = (r * r - 1.0f) * inv * rp.resp_dscale;
        } else {
          lp = y * eta - softplus_f(eta);
          deta = y - sigmoid_f(eta);
        }
      }
      if (lane < SP) { lpy += lp; ga += deta; gresp_scale += dsc; }
      // back to features: dz_ij += deta_s * beta_sj ; gbeta_sj += deta_s * z_ij
#pragma unroll
      for (int s = 0; s < SP; ++s) {
        const float de = __shfl_sync(0xffffffffu, deta, s);
        const float dzr = de * sBeta[s * 32 + lane];
        gl += dzr;
        gs = fmaf(dzr, (zs[s] - loc) / scale, gs);
        gb[s] = fmaf(de, zs[s], gb[s]);
      }
    }
    if (active) {
      // d loss = -(mean_s dLJ + dH); H = sum log scale + const, dH/dscale = 1/scale
      const float e_w = with_entropy ? 1.0f : 0.0f;
      const float out_l = -gl * invS;
      const float out_s = -(gs * invS + (first_pass ? e_w / scale : 0.0f));
      if (first_pass) { d.grad_loc[e] = out_l; d.grad_scale[e] = out_s; }
      else { d.grad_loc[e] += out_l; d.grad_scale[e] += out_s; }
      // added to every particle's log-density column (only the mean over particles is used), so
      // every pass over a particle range accumulates it
      if (with_entropy) ent += (double)(0.5f + kLogSqrt2Pi + logf(scale));
    }
  }

  // ---- CTA reduction -> one partial block per CTA -------------------------------------------
  // stage per-warp values [warp][s][40]: 0..31 gbeta per lane, 32 z2 (summed over lanes),
  // 34 (slot of s = 0) the feature residual sum, 35-37 per-particle response sums, 38 rows, 39 entropy
  float* mine = sOut + (size_t)warp * SP * 40;
#pragma unroll
  for (int s = 0; s < SP; ++s) {
    mine[s * 40 + lane] = gb[s];
    const float a = warp_sum(z2[s]);
    if (lane == 0) mine[s * 40 + 32] = a;
  }
  {
    const float c = warp_sum(f2);
    if (lane == 0) mine[34] = c;
  }
  if (lane < SP) { mine[lane * 40 + 35] = lpy; mine[lane * 40 + 36] = ga; mine[lane * 40 + 37] = gresp_scale; }
  const float rows_warp = n_rows_lane;   // identical on all lanes
  if (lane == 0) mine[38] = rows_warp;
  const double ent_w = warp_sum(ent);
  if (lane == 0) mine[39] = (float)ent_w;
  __syncthreads();
  const int ncol = 1 + p + 5;
  float* out = partial + (size_t)blockIdx.x * S * ncol;
  float rows_cta = 0.f, ent_cta = 0.f, f2_cta = 0.f;
  for (int w = 0; w < kRowWarps; ++w) {
    rows_cta += sOut[(size_t)w * SP * 40 + 38];
    ent_cta += sOut[(size_t)w * SP * 40 + 39];
    f2_cta += sOut[(size_t)w * SP * 40 + 34];
  }
  for (int idx = threadIdx.x; idx < s_count * ncol; idx += kRowThreads) {
    const int s = idx / ncol, c = idx % ncol;
    const RowParticle& rp = sPar[s];
    auto total = [&](int col) { float a = 0.f; for (int w = 0; w < kRowWarps; ++w) a += sOut[((size_t)w * SP + s) * 40 + col]; return a; };
    float val = 0.f;
    const float n_elem = rows_cta * (float)p;
    if (c == 0) {
      const float Z2 = total(32), LY = total(35);
      val = -0.5f * rp.prior_inv_var * Z2 - n_elem * (logf(rp.prior_scale) + kLogSqrt2Pi) + LY;
      // constant feature scale: every particle of this launch gets an equal share of the residual sum
      if (d.feat != nullptr)
        val += -0.5f * rp.feat_inv_var * (f2_cta / (float)s_count) - n_elem * (logf(rp.feat_scale) + kLogSqrt2Pi);
      val += ent_cta;      // entropy of q(Z), the same for every particle (only the mean over s is used)
    } else if (c <= p) {
      val = total(c - 1);
    } else if (c == p + 1) {
      val = total(36);
    } else if (c == p + 2) {
      val = 0.f;                                                            // prior location is a constant
    } else if (c == p + 3) {
      val = (total(32) * rp.prior_inv_var / rp.prior_scale - n_elem / rp.prior_scale) * rp.prior_dscale;
    } else if (c == p + 4) {
      val = 0.f;                                                            // feature scale is a constant
    } else {
      val = total(37);
    }
    out[(size_t)(s_begin + s) * ncol + c] = val;
  }
  if (bad) atomicOr(status, bad);
  if (bad_value) atomicOr(status, MNF_ST_BAD_VALUE);
}

template <int SP>
inline size_t rowlatent_smem_bytes() {
  return sizeof(float) * SP * 32 + sizeof(RowParticle) * SP + sizeof(float) * kRowWarps * SP * 40;
}

}  // namespace mnf
