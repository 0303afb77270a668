// Shared device helpers of the mininf_b200 ELBO engine: error plumbing, warp reductions, special
// functions, Philox4x32-10, and the per-family log-density + partial-derivative kernels.
//
// The log-density formulas restate torch.distributions (the reference's numeric engine,
// SURVEY.md §8a row a9) in fp32, file:line cited per family below
// (TORCH = site-packages/torch).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mininf_b200.h"

namespace mnf {

constexpr float kLogSqrt2Pi = 0.91893853320467274178f;  // log(sqrt(2*pi)), TORCH normal.py:99-103
constexpr float kFloatEps = 1.1920928955078125e-07f;    // torch.finfo(float32).eps, utils.py:101
constexpr float kFloatTiny = 1.17549435e-38f;           // torch.finfo(float32).tiny, gamma.py:85

// ---------------------------------------------------------------------------------------------
// host-side error text (thread local, returned by mnf_last_error)
// ---------------------------------------------------------------------------------------------
inline thread_local char g_last_error[512] = "";

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_last_error, sizeof(g_last_error), fmt, a, b);
  return code;
}

// kernels enqueued by this thread so far (mnf_plan_launches reports the count of one step)
inline thread_local int g_launches = 0;

// after every kernel launch: count it and surface a launch error
#define MNF_LAUNCH_CHECK()                                                                    \
  do {                                                                                        \
    ++::mnf::g_launches;                                                                      \
    MNF_CUDA_CHECK(cudaGetLastError());                                                       \
  } while (0)

#define MNF_CUDA_CHECK(expr)                                                                  \
  do {                                                                                        \
    cudaError_t err__ = (expr);                                                               \
    if (err__ != cudaSuccess) return ::mnf::fail(MNF_E_CUDA, "%s failed: %s", #expr,          \
                                                 cudaGetErrorString(err__));                  \
  } while (0)

// mnf_svi_step: the row latent's own parameters are trained inside the sweep (README.md:63-69's
// optimizer.step() and ParameterizedDistribution's exp transform, mininf/nn.py:88-96, per element).
// `enabled == 0` leaves the kernel as a pure gradient evaluation.
struct RowAdam {
  int enabled;
  float lr, beta1, beta2, eps;
  const int64_t* step;          // device: updates applied so far (the tail kernel increments it)
  float* loc_rw;                // unconstrained location (identity transform)
  float* raw_scale;             // unconstrained scale, scale = exp(raw)
  float* m_loc;
  float* v_loc;
  float* m_scale;
  float* v_scale;
};

// ---------------------------------------------------------------------------------------------
// warp / block reductions (fixed butterfly order => run-to-run deterministic)
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------
// special functions
// ---------------------------------------------------------------------------------------------
// digamma / trigamma by upward recurrence to x >= 6 and the asymptotic series. Evaluated in fp64
// when called from the O(D) kernels (entropy, implicit gradients), fp32-accurate either way.
static __device__ __noinline__ double digamma_d(double x) {
  if (x <= 0.0 && floor(x) == x) return INFINITY;
  double r = 0.0;
  if (x < 0.0) {  // reflection
    const double pi = 3.14159265358979323846;
    r = -pi / tan(pi * x);
    x = 1.0 - x;
  }
  while (x < 6.0) {
    r -= 1.0 / x;
    x += 1.0;
  }
  const double f = 1.0 / (x * x);
  const double t = f * (-1.0 / 12.0 + f * (1.0 / 120.0 + f * (-1.0 / 252.0 + f * (1.0 / 240.0 +
                   f * (-1.0 / 132.0 + f * (691.0 / 32760.0 + f * (-1.0 / 12.0)))))));
  return r + log(x) - 0.5 / x + t;
}

static __device__ __noinline__ double trigamma_d(double x) {
  double r = 0.0;
  while (x < 6.0) {
    r += 1.0 / (x * x);
    x += 1.0;
  }
  const double f = 1.0 / (x * x);
  // 1/x + 1/(2x^2) + sum B_2k / x^(2k+1)
  const double t = 1.0 / x * (1.0 + 0.5 / x + f * (1.0 / 6.0 + f * (-1.0 / 30.0 + f * (1.0 / 42.0 +
                   f * (-1.0 / 30.0 + f * (5.0 / 66.0))))));
  return r + t;
}

__device__ __forceinline__ float digamma_f(float x) { return (float)digamma_d((double)x); }

// torch.xlogy: 0 where x == 0 (even if y is 0 / inf), NaN-propagating in y.
__device__ __forceinline__ float xlogy(float x, float y) {
  if (y != y) return y;
  return x == 0.0f ? 0.0f : x * logf(y);
}

// softplus(x) = log(1 + exp(x)) in the binary_cross_entropy_with_logits form
// max(x,0) + log1p(exp(-|x|))  (ATen Loss.cpp, called from TORCH bernoulli.py:124).
__device__ __forceinline__ float softplus_f(float x) {
  return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x)));
}

__device__ __forceinline__ float sigmoid_f(float x) {
  // exp(-|x|) form: no overflow, symmetric accuracy
  const float e = expf(-fabsf(x));
  const float s = 1.0f / (1.0f + e);
  return x >= 0.0f ? s : e * s;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), the counter-based generator torch's CUDA generator also
// uses (TORCH include/ATen/core/PhiloxRNGEngine.h). key = seed, counter = (offset, index).
// ---------------------------------------------------------------------------------------------
// Stream tags: the top byte of the 64-bit index keeps the consumers of one (seed, call index)
// apart, so no two draws of a step can share a counter (packed Normal draws use index = s*D + d,
// Gamma draws (index << 8) | round, row latents (element << 8) | particle group: without a tag
// those ranges overlap once S*D >= 256).
constexpr uint64_t kPhiloxNormal = (uint64_t)1 << 56;
constexpr uint64_t kPhiloxGamma = (uint64_t)2 << 56;
constexpr uint64_t kPhiloxBeta0 = (uint64_t)3 << 56;
constexpr uint64_t kPhiloxRowLatent = (uint64_t)4 << 56;
constexpr uint64_t kPhiloxPredict = (uint64_t)5 << 56;        // posterior-predictive draws (predict.cuh)
constexpr uint64_t kPhiloxPredictGamma = (uint64_t)4 << 48;   // | into the gamma sampler's index (tag 2 stays on top)

struct Philox {
  uint32_t c[4];
  uint32_t k[2];
  __device__ __forceinline__ Philox(uint64_t seed, uint64_t offset, uint64_t index) {
    k[0] = (uint32_t)seed;
    k[1] = (uint32_t)(seed >> 32);
    c[0] = (uint32_t)offset;
    c[1] = (uint32_t)(offset >> 32);
    c[2] = (uint32_t)index;
    c[3] = (uint32_t)(index >> 32);
  }
  __device__ __forceinline__ void round_once() {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k[0];
    const uint32_t n2 = hi0 ^ c[3] ^ k[1];
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
  }
  // ROUNDS = 10 is the standard generator (and torch's); 7 is the smallest round count of Salmon et
  // al. (SC'11, table 2) that passes BigCrush ("Crush-resistant"), offered by Random123 as
  // philox4x32_7. The row-latent sweep draws S * N * p = 1e10 normals per step and is bound by
  // instruction issue, a fifth of it Philox rounds: it uses 7 (rowlatent.cuh); everything else 10.
  template <int ROUNDS = 10>
  __device__ __forceinline__ uint4 next() {
    uint32_t k0 = k[0], k1 = k[1];
#pragma unroll
    for (int r = 0; r < ROUNDS; ++r) {
      round_once();
      k[0] += 0x9E3779B9u;
      k[1] += 0xBB67AE85u;
    }
    k[0] = k0; k[1] = k1;
    return make_uint4(c[0], c[1], c[2], c[3]);
  }
};

// Two standard normals from two 32-bit words (Box-Muller on (0,1] x [0,1) uniforms).
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
  const float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0, 1]
  const float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);           // [0, 1)
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  return make_float2(r * c, r * s);
}

// ---------------------------------------------------------------------------------------------
// log-density and partial derivatives, one (value, parameters) pair at a time
// ---------------------------------------------------------------------------------------------
// log(y!) = lgamma(y + 1) of a Poisson count: integers below 64 from a constant-memory table,
// y >= 8 from the Stirling series (truncation < 3e-8 absolute), anything else (invalid data,
// flagged separately) from lgammaf. As accurate as lgammaf at a fraction of its instruction count.
static __constant__ float kLogFactorial[64] = {
    0.0f, 0.0f, 0.693147181f, 1.79175947f,
    3.17805383f, 4.78749174f, 6.57925121f, 8.52516136f,
    10.6046029f, 12.8018275f, 15.1044126f, 17.5023078f,
    19.9872145f, 22.5521639f, 25.1912212f, 27.8992714f,
    30.6718601f, 33.5050735f, 36.3954452f, 39.3398842f,
    42.3356165f, 45.3801389f, 48.4711814f, 51.6066756f,
    54.7847294f, 58.0036052f, 61.2617018f, 64.5575386f,
    67.8897431f, 71.257039f, 74.6582363f, 78.0922236f,
    81.5579595f, 85.054467f, 88.5808275f, 92.1361756f,
    95.7196945f, 99.3306125f, 102.968199f, 106.63176f,
    110.32064f, 114.034212f, 117.771881f, 121.533082f,
    125.317271f, 129.123934f, 132.952575f, 136.802723f,
    140.673924f, 144.565744f, 148.477767f, 152.409593f,
    156.360836f, 160.331128f, 164.320112f, 168.327445f,
    172.352797f, 176.395848f, 180.456291f, 184.533829f,
    188.628173f, 192.739047f, 196.866182f, 201.009316f,
};
__device__ __forceinline__ float log_factorial(float y) {
  if (y >= 0.0f && y < 64.0f && y == floorf(y)) return kLogFactorial[(int)y];
  if (y >= 8.0f) {
    const float x = y + 1.0f, r = __fdividef(1.0f, x), r2 = r * r;
    return (x - 0.5f) * logf(x) - x + 0.91893853320467274f + r * (0.083333333f - r2 * (0.0027777778f - r2 * 0.00079365079f));
  }
  return lgammaf(y + 1.0f);
}

// support of an observed value (torch.distributions' validate_args / mininf/core.py:183)
__device__ __forceinline__ bool in_support(int family, float v) {
  switch (family) {
    case MNF_NORMAL: return v == v;
    case MNF_GAMMA: return v >= 0.0f;
    case MNF_BETA: return v >= 0.0f && v <= 1.0f;
    case MNF_POISSON: return v >= 0.0f && floorf(v) == v;
    default: return v == 0.0f || v == 1.0f;
  }
}

struct Dens {
  float lp;  // log p(v | p0, p1)
  float dv;  // d lp / d v
  float d0;  // d lp / d p0
  float d1;  // d lp / d p1
  bool bad_param;
  bool bad_value;
};

// `need_grad` lets callers skip the digamma evaluations when the parameters are constants.
__device__ inline Dens density(int family, float v, float p0, float p1, bool need_grad) {
  Dens o;
  o.lp = 0.f; o.dv = 0.f; o.d0 = 0.f; o.d1 = 0.f; o.bad_param = false; o.bad_value = false;
  switch (family) {
    case MNF_NORMAL: {  // TORCH distributions/normal.py:87-103
      o.bad_param = !(p1 > 0.0f) || p0 != p0;
      o.bad_value = v != v;
      const float inv = 1.0f / p1;
      const float r = (v - p0) * inv;
      o.lp = -0.5f * r * r - logf(p1) - kLogSqrt2Pi;
      o.d0 = r * inv;
      o.dv = -o.d0;
      o.d1 = (r * r - 1.0f) * inv;
      break;
    }
    case MNF_GAMMA: {  // TORCH distributions/gamma.py:89-98
      o.bad_param = !(p0 > 0.0f) || !(p1 > 0.0f);
      o.bad_value = !(v >= 0.0f);
      o.lp = xlogy(p0, p1) + xlogy(p0 - 1.0f, v) - p1 * v - lgammaf(p0);
      o.dv = (p0 - 1.0f) / v - p1;
      if (need_grad) {
        o.d0 = logf(p1) + logf(v) - digamma_f(p0);
        o.d1 = p0 / p1 - v;
      }
      break;
    }
    case MNF_BETA: {  // TORCH distributions/beta.py:87-91 -> dirichlet.py:90-97 (p0 = concentration1)
      o.bad_param = !(p0 > 0.0f) || !(p1 > 0.0f);
      o.bad_value = !(v >= 0.0f && v <= 1.0f);
      o.lp = xlogy(p0 - 1.0f, v) + xlogy(p1 - 1.0f, 1.0f - v) + lgammaf(p0 + p1) - lgammaf(p0) -
             lgammaf(p1);
      o.dv = (p0 - 1.0f) / v - (p1 - 1.0f) / (1.0f - v);
      if (need_grad) {
        const float dt = digamma_f(p0 + p1);
        o.d0 = logf(v) + dt - digamma_f(p0);
        o.d1 = log1pf(-v) + dt - digamma_f(p1);
      }
      break;
    }
    case MNF_BERNOULLI_PROBS: {  // TORCH bernoulli.py:121-125, utils.py:101-137 (clamp + logit)
      o.bad_param = !(p0 >= 0.0f && p0 <= 1.0f);
      o.bad_value = !(v == 0.0f || v == 1.0f);
      const float pc = fminf(fmaxf(p0, kFloatEps), 1.0f - kFloatEps);
      const float logits = logf(pc) - log1pf(-pc);
      o.lp = v * logits - softplus_f(logits);
      // d lp/d logits = v - sigmoid(logits) = v - pc ; d logits/d p = 1/(pc (1-pc)) inside the clamp
      const bool inside = p0 >= kFloatEps && p0 <= 1.0f - kFloatEps;
      o.d0 = inside ? (v - pc) / (pc * (1.0f - pc)) : 0.0f;
      break;
    }
    case MNF_BERNOULLI_LOGITS: {  // TORCH bernoulli.py:121-125
      o.bad_param = p0 != p0;
      o.bad_value = !(v == 0.0f || v == 1.0f);
      o.lp = v * p0 - softplus_f(p0);
      o.d0 = v - sigmoid_f(p0);
      break;
    }
    case MNF_POISSON: {  // TORCH poisson.py:75-79
      o.bad_param = !(p0 >= 0.0f);
      o.bad_value = !(v >= 0.0f) || floorf(v) != v;
      o.lp = xlogy(v, p0) - p0 - lgammaf(v + 1.0f);
      o.d0 = v / p0 - 1.0f;
      if (v == 0.0f) o.d0 = -1.0f;
      break;
    }
    default:
      o.bad_param = true;
  }
  return o;
}

// One scalar link T(A + B x) for particle row `zs` (pointer to z[s][0]) and element i.
struct LinkVal {
  float value;  // T(u)
  float du;     // dT/du at u
  float x;      // covariate value (1 when absent)
};

__device__ __forceinline__ LinkVal eval_link(const mnf_link_t& L, const float* zs, int64_t i) {
  const float A = L.a_const + (L.a_lat < 0 ? 0.0f : zs[L.a_lat + (int64_t)L.a_stride * i]);
  const float B = L.b_const + (L.b_lat < 0 ? 0.0f : zs[L.b_lat + (int64_t)L.b_stride * i]);
  const float x = L.x == nullptr ? 1.0f : __ldg(L.x + (int64_t)L.x_stride * i);
  const float u = fmaf(B, x, A);
  LinkVal o;
  o.x = x;
  if (L.transform == MNF_T_EXP) {
    o.value = expf(u);
    o.du = o.value;
  } else {
    o.value = u;
    o.du = 1.0f;
  }
  return o;
}

__device__ __forceinline__ bool link_has_latent(const mnf_link_t& L) {
  return L.a_lat >= 0 || L.b_lat >= 0;
}

}  // namespace mnf
