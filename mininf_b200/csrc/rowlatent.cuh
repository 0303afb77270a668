// Row-latent sweep: a per-observation latent matrix Z [n][p] (p <= 32) with a mean-field Normal
// approximation, and the prior / feature / response sites that touch it, for all particles in one
// pass (config C4: examples/regression-with-feature-uncertainty.md:28-38 widened to p features).
//
// Replaces, per step: Normal.rsample of n*p values per particle (TORCH normal.py:82-85), the
// element-wise log_prob chains of three sites (mininf/core.py:241), `z @ slope` and its backward,
// the entropy of q(Z) (TORCH normal.py:114-115) and autograd's n*p-sized gradient passes.
//
// Mapping: a warp owns a row, lane j owns feature j. Per row a lane keeps loc, scale and feature
// in registers and loops over the particles in groups of four (one Philox4x32-7 call -> two
// Box-Muller pairs). The kernel is instruction-bound, so the per-(element, particle)
// work is cut to what cannot be shared:
//   * with z = loc + scale eps the feature site needs only E1 = sum_s eps, E2 = sum_s eps^2 per
//     element (its scale is a constant): sum_s (x - z_s), sum_s (x - z_s) eps_s and
//     sum_s (x - z_s)^2 are polynomials in them;
//   * the prior site (scale sigma_s is a per-particle latent) needs W1 = sum_s eps_s / sigma_s^2,
//     W2 = sum_s eps_s^2 / sigma_s^2 per element and sum_ij (z - m)^2 per particle;
//   * the dot products Z_i . beta_s are reduced AND transposed in one 31-shuffle butterfly so that
//     lane s ends up with particle s's linear predictor: one exp per lane and row;
//   * d eta is handed back through a shared-memory slot read as 128-bit broadcasts; beta lives in
//     shared memory as [lane][particle] (padded) so four particles come with one LDS.128.
// Per-(particle, lane) statistics live in registers (SP is a template parameter), so the hot loop
// has no atomics. Algorithmic bytes per row: p * (loc 4 + scale 4 + feature 4 + two gradients 8)
// + response 4.
#pragma once

#include "common.cuh"

namespace mnf {

#ifndef MNF_ROWLATENT_PHILOX_ROUNDS
#define MNF_ROWLATENT_PHILOX_ROUNDS 7
#endif
constexpr int kRowLatentPhiloxRounds = MNF_ROWLATENT_PHILOX_ROUNDS;   // see common.cuh::Philox::next


constexpr int kRowThreads = 128;
constexpr int kRowWarps = kRowThreads / 32;

struct RowParticle {   // per-particle scalars of the row-latent sites, staged in shared memory
  float prior_inv_var, prior_scale, prior_dscale;  // Normal prior of Z
  float icpt, resp_scale, resp_dscale;
};

// Two standard normals from two 32-bit words: Box-Muller on the MUFU pipe (lg2, sqrt, sin, cos
// approximations; absolute error of the draws ~1e-6). The uniforms are built by bit insertion
// (23 bits each, no integer->float conversion): u1 in (0, 1], the angle in [-pi, pi).
__device__ __forceinline__ float2 box_muller_fast(uint32_t a, uint32_t b) {
  const float f1 = __uint_as_float(0x3f800000u | (a >> 9));    // [1, 2)
  const float f2 = __uint_as_float(0x3f800000u | (b >> 9));
  const float u1 = 2.0f - f1;                                   // (0, 1]
  const float theta = fmaf(f2, 6.2831853071795865f, -9.4247779607693797f);
  float lg, r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u1));          // u1 >= 2^-23: never denormal
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * lg));   // sqrt(-2 ln u1)
  return make_float2(r * __cosf(theta), r * __sinf(theta));
}

// exp(u) as ex2(u * log2 e) with the product carried in two terms (relative error ~2 ulp)
__device__ __forceinline__ float exp_fast(float u) {
  const float t = fmaf(u, 1.925963033500011e-8f, u * 1.4426950216293335f);
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
  return r;
}

template <int SP>
struct RowSmem {
  static constexpr int kBetaStride = SP + 4;             // floats; 16-byte rows, LDS.128 conflict-free
  float beta[32 * kBetaStride];                          // [lane][particle]
  float ipv[SP];                                         // prior 1/sigma_s^2 (0 for dead particles)
  float deta[kRowWarps][SP];                             // d log p(y) / d eta of the row in flight
  RowParticle par[SP];
  float out[kRowWarps][SP][40];                          // final reduction scratch
};

// partial layout per CTA: [S][ncol], ncol = 1 + p + 5:
//   0 log-density (+ entropy share), 1..p beta gradient, p+1 intercept, p+2 prior loc (du),
//   p+3 prior scale (du), p+4 feature scale (du), p+5 response scale (du)
// FULL: every particle slot of this launch is in use (s_count == SP): no masking code.
template <int SP, bool FULL>
__global__ void __launch_bounds__(kRowThreads)
rowlatent_kernel(mnf_rowlatent_t d, const float* __restrict__ z, int S, int D, int s_begin,
                 int first_pass, uint64_t seed, uint64_t offset, const uint64_t* __restrict__ offset_dev,
                 int with_entropy, float* __restrict__ partial, uint32_t* __restrict__ status, RowAdam adam) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RowSmem<SP>& sm = *reinterpret_cast<RowSmem<SP>*>(smem_raw);
  constexpr int BS = RowSmem<SP>::kBetaStride;
  if (offset_dev != nullptr) offset += *offset_dev;   // device-side call index (CUDA-graph replays)
  const int p = d.p;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool has_feat = d.feat != nullptr, has_resp = d.resp != nullptr;

  const int s_count = FULL ? SP : min(SP, S - s_begin);   // particles handled by this launch
  for (int i = threadIdx.x; i < SP * 32; i += kRowThreads) {
    const int s = i >> 5, j = i & 31;
    sm.beta[j * BS + s] = (s < s_count && j < p && has_resp) ? z[(int64_t)(s_begin + s) * D + d.beta_lat + j] : 0.0f;
  }
  uint32_t bad = 0;
  // prior location and feature scale are constants in this build (checked on the host)
  const float prior_loc = d.prior_loc.a_const + d.prior_loc.b_const;
  float feat_scale = 1.0f, feat_inv_var = 0.0f;
  if (has_feat) {
    const float u = d.feat_scale.a_const + d.feat_scale.b_const;
    feat_scale = d.feat_scale.transform == MNF_T_EXP ? expf(u) : u;
    feat_inv_var = 1.0f / (feat_scale * feat_scale);
    if (!(feat_scale > 0.0f)) bad |= MNF_ST_BAD_PARAM;
  }
  for (int s = threadIdx.x; s < SP; s += kRowThreads) {
    RowParticle rp;
    rp.prior_inv_var = 0.f; rp.prior_scale = 1.f; rp.prior_dscale = 0.f;
    rp.icpt = 0.f; rp.resp_scale = 1.f; rp.resp_dscale = 0.f;
    if (s < s_count) {
      const float* zs = z + (int64_t)(s_begin + s) * D;
      const LinkVal ps = eval_link(d.prior_scale, zs, 0);
      rp.prior_scale = ps.value; rp.prior_dscale = ps.du;
      rp.prior_inv_var = 1.0f / (ps.value * ps.value);
      if (!(ps.value > 0.0f)) bad |= MNF_ST_BAD_PARAM;
      if (has_resp) {
        rp.icpt = d.icpt_const + (d.icpt_lat >= 0 ? zs[d.icpt_lat] : 0.0f);
        if (d.resp_family == MNF_NORMAL) {
          const LinkVal rs = eval_link(d.resp_scale, zs, 0);
          rp.resp_scale = rs.value; rp.resp_dscale = rs.du;
          if (!(rs.value > 0.0f)) bad |= MNF_ST_BAD_PARAM;
        }
      }
    }
    sm.par[s] = rp;
    sm.ipv[s] = rp.prior_inv_var;
  }
  __syncthreads();
  float ipv_sum = 0.f;     // sum over the live particles of 1/sigma_s^2
  for (int s = 0; s < SP; ++s) ipv_sum += sm.ipv[s];

  // per-(particle, lane) statistics, registers
  float z2[SP], gb[SP];      // sum (z - prior_loc)^2 ; sum dlp/deta * z
#pragma unroll
  for (int s = 0; s < SP; ++s) { z2[s] = 0.f; gb[s] = 0.f; }
  // per-particle statistics owned by lane s (after the butterfly lane s holds particle s)
  float ga = 0.f, lpy = 0.f, gresp_scale = 0.f;
  float f2 = 0.f;            // sum over particles and rows of (x - z)^2
  double ent = 0.0;          // entropy of q(Z)
  float n_rows_lane = 0.f;
  bool bad_value = false;
  const float invS = 1.0f / (float)S;
  const float n_live = (float)s_count;
  const bool active = lane < p;
  const float4* beta4 = reinterpret_cast<const float4*>(sm.beta + lane * BS);
  const float4* ipv4 = reinterpret_cast<const float4*>(sm.ipv);
  float* my_deta = sm.deta[warp];
  const int sl = lane & (SP - 1);
  const RowParticle rp_l = sm.par[sl];      // the particle this lane owns after the butterfly
  const bool lane_live = sl < s_count;

  // fused optimiser: bias corrections of this step (torch.optim.Adam), the scale read through exp
  float adam_step_size = 0.f, adam_inv_bc2_sqrt = 0.f;
  if (adam.enabled) {
    const double t = (double)(*adam.step + 1);
    adam_step_size = (float)((double)adam.lr / (1.0 - pow((double)adam.beta1, t)));
    adam_inv_bc2_sqrt = (float)(1.0 / sqrt(1.0 - pow((double)adam.beta2, t)));
  }
  const float* scale_src = adam.enabled ? adam.raw_scale : d.scale;
  const int64_t warp_global = (int64_t)blockIdx.x * kRowWarps + warp;
  const int64_t warps_total = (int64_t)gridDim.x * kRowWarps;
  // the next row's operands are fetched while the current row is being worked on
  float n_loc = prior_loc, n_scale = 1.f, n_x = 0.f, n_y = 0.f;
  if (warp_global < d.n_rows) {
    const int64_t e0 = warp_global * p + lane;
    if (active) { n_loc = d.loc[e0]; n_scale = scale_src[e0]; }
    if (active && has_feat) n_x = __ldg(d.feat + e0);
    if (has_resp) n_y = __ldg(d.resp + warp_global);
  }
  for (int64_t row = warp_global; row < d.n_rows; row += warps_total) {
    const int64_t e = row * p + lane;
    // inactive lanes (feature index >= p) are inert: z == prior_loc == x, scale 0, beta 0
    const float loc = active ? n_loc : prior_loc;
    const float raw_scale = n_scale;                       // unconstrained value (fused optimiser only)
    const float scale_raw = active ? (adam.enabled ? expf(n_scale) : n_scale) : 1.f;
    const float scale = active ? scale_raw : 0.f;
    const float x = (active && has_feat) ? n_x : loc;
    const float y = n_y;
    {
      const int64_t row_n = row + warps_total;
      if (row_n < d.n_rows) {
        const int64_t en = row_n * p + lane;
        // plain loads: with the fused optimiser these arrays are written by this kernel (every
        // element exactly once, by the thread that read it)
        if (active) { n_loc = d.loc[en]; n_scale = scale_src[en]; }
        if (active && has_feat) n_x = __ldg(d.feat + en);
        if (has_resp) n_y = __ldg(d.resp + row_n);
      }
    }
    // fused optimiser: this element's Adam moments are requested now and consumed after the
    // particle loop, ~1500 instructions later
    float am_l = 0.f, av_l = 0.f, am_s = 0.f, av_s = 0.f;
    if (adam.enabled && active) {
      am_l = adam.m_loc[e]; av_l = adam.v_loc[e]; am_s = adam.m_scale[e]; av_s = adam.v_scale[e];
    }
    if (!(scale_raw > 0.0f)) bad |= MNF_ST_BAD_PARAM;
    if (y != y || x != x) bad_value = true;
    float eps_r[SP], v[SP];
    float E1 = 0.f, E2 = 0.f, W1 = 0.f, W2 = 0.f;
    n_rows_lane += 1.f;
    const float dloc = loc - prior_loc;
    // ---- draws: one uniform branch per row, so the independent Philox chains of the row share a
    // basic block and the scheduler can interleave them ---------------------------------------
    if (d.eps != nullptr) {
#pragma unroll
      for (int s = 0; s < SP; ++s) {
        const int sg = s_begin + s;
        eps_r[s] = (active && sg < S) ? __ldg(d.eps + ((int64_t)sg * d.n_rows + row) * p + lane) : 0.f;
      }
    } else {
#pragma unroll
      for (int q = 0; q < SP / 4; ++q) {
        Philox rng(seed, offset, kPhiloxRowLatent | ((uint64_t)e << 8) | (uint64_t)((s_begin >> 2) + q));
        const uint4 r = rng.next<kRowLatentPhiloxRounds>();
        const float2 n0 = box_muller_fast(r.x, r.y), n1 = box_muller_fast(r.z, r.w);
        eps_r[4 * q + 0] = n0.x; eps_r[4 * q + 1] = n0.y; eps_r[4 * q + 2] = n1.x; eps_r[4 * q + 3] = n1.y;
      }
      if (!FULL) {
#pragma unroll
        for (int s = 0; s < SP; ++s)
          if (s >= s_count) eps_r[s] = 0.f;
      }
    }
    // ---- the element-wise sites ----------------------------------------------------------------
#pragma unroll
    for (int q = 0; q < SP / 4; ++q) {
      const float4 b4 = beta4[q], i4 = ipv4[q];
      const float bq[4] = {b4.x, b4.y, b4.z, b4.w}, iq[4] = {i4.x, i4.y, i4.z, i4.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int s = 4 * q + t;
        const float eps = eps_r[s];
        const float dzp = fmaf(eps, scale, dloc);          // z - prior_loc
        z2[s] = fmaf(dzp, dzp, z2[s]);
        const float tw = iq[t] * eps;
        W1 += tw;
        W2 = fmaf(tw, eps, W2);
        E1 += eps;
        E2 = fmaf(eps, eps, E2);
        v[s] = fmaf(eps, scale, loc) * bq[t];
      }
    }
    // prior: d/dz = -(z - m)/sigma_s^2 ; features: d/dz = (x - z)/ns^2, summed over particles
    //   sum_s (z_s - m)/sigma_s^2       = dloc * ipv_sum + scale * W1
    //   sum_s (z_s - m) eps_s/sigma_s^2 = dloc * W1 + scale * W2
    //   sum_s (x - z_s)                 = n dx - scale E1          dx = x - loc
    //   sum_s (x - z_s) eps_s           = dx E1 - scale E2
    //   sum_s (x - z_s)^2               = n dx^2 - 2 dx scale E1 + scale^2 E2
    float gl = -fmaf(dloc, ipv_sum, scale * W1);
    float gs = -fmaf(dloc, W1, scale * W2);
    if (has_feat) {
      const float dx = x - loc;
      gl = fmaf(feat_inv_var, fmaf(n_live, dx, -scale * E1), gl);
      gs = fmaf(feat_inv_var, fmaf(dx, E1, -scale * E2), gs);
      const float se = scale * E1;
      f2 += fmaf(scale * scale, E2, fmaf(n_live * dx, dx, -2.0f * dx * se));
    }
    // ---- response: transpose-reduce the dot products so lane s owns particle s -----------------
    if (has_resp) {
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
      const float eta = rp_l.icpt + v[0];
      float deta = 0.f, lp = 0.f, dsc = 0.f;
      if (lane_live) {
        if (d.resp_family == MNF_POISSON) {
          if (d.resp_transform == MNF_T_EXP) {
            const float rate = exp_fast(eta);
            lp = fmaf(y, eta, -rate) - log_factorial(y);
            deta = y - rate;
          } else {
            lp = xlogy(y, eta) - eta - log_factorial(y);
            deta = y == 0.f ? -1.f : y / eta - 1.0f;
          }
        } else if (d.resp_family == MNF_NORMAL) {
          const float inv = 1.0f / rp_l.resp_scale;
          const float r = (y - eta) * inv;
          lp = -0.5f * r * r - logf(rp_l.resp_scale) - kLogSqrt2Pi;
          deta = r * inv;
          dsc = (r * r - 1.0f) * inv * rp_l.resp_dscale;
        } else {
          lp = y * eta - softplus_f(eta);
          deta = y - sigmoid_f(eta);
        }
      }
      if (lane < SP) { lpy += lp; ga += deta; gresp_scale += dsc; my_deta[lane] = deta; }
      __syncwarp();
      // back to features: dz_ij += deta_s * beta_sj ; gbeta_sj += deta_s * z_ij
      const float4* de4 = reinterpret_cast<const float4*>(my_deta);
#pragma unroll
      for (int q = 0; q < SP / 4; ++q) {
        const float4 d4 = de4[q], b4 = beta4[q];
        const float dq[4] = {d4.x, d4.y, d4.z, d4.w}, bq[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const int s = 4 * q + t;
          const float dzr = dq[t] * bq[t];
          gl += dzr;
          gs = fmaf(dzr, eps_r[s], gs);
          gb[s] = fmaf(dq[t], fmaf(eps_r[s], scale, loc), gb[s]);
        }
      }
      __syncwarp();   // the slot is rewritten by the next row
    }
    if (active) {
      // d loss = -(mean_s dLJ + dH); H = sum log scale + const, dH/dscale = 1/scale
      const float e_w = with_entropy ? 1.0f : 0.0f;
      const float out_l = -gl * invS;
      const float out_s = -(gs * invS + (first_pass ? __fdividef(e_w, scale) : 0.0f));
      if (adam.enabled) {
        // chain rule through scale = exp(raw): d/d raw = d/d scale * scale; then Adam, in place
        const float g_l = out_l, g_s = out_s * scale;
        const float m_l = adam.beta1 * am_l + (1.0f - adam.beta1) * g_l;
        const float v_l = adam.beta2 * av_l + (1.0f - adam.beta2) * g_l * g_l;
        const float m_s = adam.beta1 * am_s + (1.0f - adam.beta1) * g_s;
        const float v_s = adam.beta2 * av_s + (1.0f - adam.beta2) * g_s * g_s;
        adam.m_loc[e] = m_l; adam.v_loc[e] = v_l; adam.m_scale[e] = m_s; adam.v_scale[e] = v_s;
        adam.loc_rw[e] = loc - adam_step_size * __fdividef(m_l, fmaf(sqrtf(v_l), adam_inv_bc2_sqrt, adam.eps));
        adam.raw_scale[e] = raw_scale - adam_step_size * __fdividef(m_s, fmaf(sqrtf(v_s), adam_inv_bc2_sqrt, adam.eps));
      } else if (first_pass) { d.grad_loc[e] = out_l; d.grad_scale[e] = out_s; }
      else { d.grad_loc[e] += out_l; d.grad_scale[e] += out_s; }
      // added to every particle's log-density column (only the mean over particles is used), so
      // every pass over a particle range accumulates it
      if (with_entropy) ent += (double)(0.5f + kLogSqrt2Pi + __logf(scale));
    }
  }

  // ---- CTA reduction -> one partial block per CTA -------------------------------------------
  // stage per-warp values [warp][s][40]: 0..31 gbeta per lane, 32 z2 (summed over lanes),
  // 34 (slot of s = 0) the feature residual sum, 35-37 per-particle response sums, 38 rows, 39 entropy
  float* mine = &sm.out[warp][0][0];
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
    rows_cta += sm.out[w][0][38];
    ent_cta += sm.out[w][0][39];
    f2_cta += sm.out[w][0][34];
  }
  for (int idx = threadIdx.x; idx < s_count * ncol; idx += kRowThreads) {
    const int s = idx / ncol, c = idx % ncol;
    const RowParticle& rp = sm.par[s];
    auto total = [&](int col) { float a = 0.f; for (int w = 0; w < kRowWarps; ++w) a += sm.out[w][s][col]; return a; };
    float val = 0.f;
    const float n_elem = rows_cta * (float)p;
    if (c == 0) {
      const float Z2 = total(32), LY = total(35);
      val = -0.5f * rp.prior_inv_var * Z2 - n_elem * (logf(rp.prior_scale) + kLogSqrt2Pi) + LY;
      // constant feature scale: every particle of this launch gets an equal share of the residual sum
      if (has_feat)
        val += -0.5f * feat_inv_var * (f2_cta / (float)s_count) - n_elem * (logf(feat_scale) + kLogSqrt2Pi);
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
  return sizeof(RowSmem<SP>);
}

}  // namespace mnf
