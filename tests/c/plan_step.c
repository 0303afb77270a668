/*
 * A plain C consumer of include/mininf_b200.h: one ELBO + gradient evaluation and one fused SVI
 * step of a Bayesian linear regression (tests/test_mininf.py:7-12 of the reference with a unit
 * noise scale) from host-built tables - no Python, no torch. The loss of
 * mininf/nn.py:212-228 is recomputed here in double precision with the noise the engine drew.
 *
 *   theta ~ Normal(0, 1)^p;  y ~ Normal(X theta, 1);  q(theta) = Normal(loc, scale)
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mininf_b200.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define CHECK_MNF(x) do { int r_ = (x); if (r_ != 0) { printf("%s -> %d: %s\n", #x, r_, mnf_last_error()); return 1; } } while (0)

enum { N = 4096, P = 64, S = 8, D = P };

int main(void) {
  static float X[N * P], y[N], loc[P], scale[P], raw[2 * D];
  unsigned int lcg = 12345u;
  for (int i = 0; i < N * P; ++i) { lcg = lcg * 1664525u + 1013904223u; X[i] = ((float)(lcg >> 8) / 8388608.0f) - 1.0f; }
  for (int i = 0; i < N; ++i) {
    double eta = 0.0;
    for (int j = 0; j < P; ++j) eta += X[i * P + j] * (0.05 * (j % 7 - 3));
    lcg = lcg * 1664525u + 1013904223u;
    y[i] = (float)eta + 0.5f * (((float)(lcg >> 8) / 8388608.0f) - 1.0f);
  }
  for (int j = 0; j < P; ++j) { loc[j] = 0.01f * (float)(j % 5); scale[j] = 0.1f; raw[j] = loc[j]; raw[D + j] = logf(scale[j]); }

  float *dX, *dy, *dP, *dz, *dnoise, *dout, *draw, *dm, *dv;
  double* dacc;
  uint32_t* dstatus;
  uint8_t* dcode;
  int64_t* dsteps;
  CHECK_CUDA(cudaMalloc((void**)&dX, sizeof(X)));
  CHECK_CUDA(cudaMalloc((void**)&dy, sizeof(y)));
  CHECK_CUDA(cudaMalloc((void**)&dP, 2 * D * sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&dz, S * D * sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&dnoise, S * D * sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&dacc, S * (D + 1) * sizeof(double)));
  CHECK_CUDA(cudaMalloc((void**)&dout, (1 + 2 * D) * sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&dstatus, sizeof(uint32_t)));
  CHECK_CUDA(cudaMalloc((void**)&draw, 2 * D * sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&dm, 2 * D * sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&dv, 2 * D * sizeof(float)));
  CHECK_CUDA(cudaMalloc((void**)&dcode, 2 * D));
  CHECK_CUDA(cudaMalloc((void**)&dsteps, sizeof(int64_t)));
  CHECK_CUDA(cudaMemcpy(dX, X, sizeof(X), cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(dy, y, sizeof(y), cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(dP, loc, sizeof(loc), cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemcpy(dP + D, scale, sizeof(scale), cudaMemcpyHostToDevice));
  CHECK_CUDA(cudaMemset(dstatus, 0, sizeof(uint32_t)));
  CHECK_CUDA(cudaMemset(dm, 0, 2 * D * sizeof(float)));
  CHECK_CUDA(cudaMemset(dv, 0, 2 * D * sizeof(float)));
  CHECK_CUDA(cudaMemset(dsteps, 0, sizeof(int64_t)));
  CHECK_CUDA(cudaMemcpy(draw, raw, sizeof(raw), cudaMemcpyHostToDevice));
  uint8_t code[2 * D];
  for (int j = 0; j < D; ++j) { code[j] = MNF_T_ID; code[D + j] = MNF_T_EXP; }
  CHECK_CUDA(cudaMemcpy(dcode, code, sizeof(code), cudaMemcpyHostToDevice));

  /* the flat tables: one latent site, one dense observed site, one prior site */
  mnf_latent_t latent = {MNF_NORMAL, P, 0, 0, dP, dP + D};
  mnf_link_t one = {1.0f, 0.0f, -1, -1, 0, 0, NULL, 0, MNF_T_ID};
  mnf_link_t zero = {0.0f, 0.0f, -1, -1, 0, 0, NULL, 0, MNF_T_ID};
  mnf_dense_site_t dense;
  memset(&dense, 0, sizeof(dense));
  dense.family = MNF_NORMAL; dense.p = P; dense.n_rows = N; dense.ldx = P; dense.X = dX; dense.y = dy;
  dense.theta_lat = 0; dense.icpt_lat = -1; dense.scale = one; dense.weight = 1.0;
  int32_t mode = MNF_DENSE_TF32;
  mnf_site_t prior;
  memset(&prior, 0, sizeof(prior));
  prior.family = MNF_NORMAL; prior.value_lat = 0; prior.numel = P; prior.scale = 1.0;
  prior.param[0] = zero; prior.param[1] = one;
  mnf_plan_desc_t desc;
  memset(&desc, 0, sizeof(desc));
  desc.n_particles = S; desc.n_latent_total = D; desc.n_latents = 1; desc.latents = &latent;
  desc.n_dense = 1; desc.dense = &dense; desc.dense_mode = &mode;
  desc.n_small_global = 1; desc.small_global = &prior; desc.device = -1;

  mnf_plan_t* plan = NULL;
  CHECK_MNF(mnf_plan_create(&desc, &plan));
  size_t ws_bytes = 0;
  CHECK_MNF(mnf_plan_workspace_bytes(plan, &ws_bytes));
  void* ws;
  CHECK_CUDA(cudaMalloc(&ws, ws_bytes));
  mnf_buffers_t buffers;
  memset(&buffers, 0, sizeof(buffers));
  buffers.z = dz; buffers.noise = dnoise; buffers.acc = dacc; buffers.out = dout; buffers.workspace = ws;
  buffers.workspace_bytes = ws_bytes; buffers.status = dstatus;

  CHECK_MNF(mnf_elbo_fwd_bwd(plan, &buffers, 1234u, 1u, MNF_STEP_ENTROPY | MNF_STEP_ALL, NULL));
  CHECK_CUDA(cudaDeviceSynchronize());
  int launches = 0;
  CHECK_MNF(mnf_plan_launches(plan, &launches));
  static float noise[S * D], out[1 + 2 * D];
  uint32_t status = 0;
  CHECK_CUDA(cudaMemcpy(noise, dnoise, sizeof(noise), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(out, dout, sizeof(out), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(&status, dstatus, sizeof(status), cudaMemcpyDeviceToHost));

  /* the same estimate in double precision: -(mean_s [log p(y | theta_s) + log p(theta_s)] + H[q]) */
  const double log_sqrt_2pi = 0.91893853320467274178;
  double total = 0.0, grad_loc[P] = {0};
  for (int s = 0; s < S; ++s) {
    double theta[P], lp = 0.0;
    for (int j = 0; j < P; ++j) {
      theta[j] = (double)loc[j] + (double)noise[s * D + j] * (double)scale[j];
      lp += -0.5 * theta[j] * theta[j] - log_sqrt_2pi;
      grad_loc[j] += -theta[j];
    }
    for (int i = 0; i < N; ++i) {
      double eta = 0.0;
      for (int j = 0; j < P; ++j) eta += (double)X[i * P + j] * theta[j];
      const double r = (double)y[i] - eta;
      lp += -0.5 * r * r - log_sqrt_2pi;
      for (int j = 0; j < P; ++j) grad_loc[j] += r * (double)X[i * P + j];
    }
    total += lp;
  }
  double entropy = 0.0;
  for (int j = 0; j < P; ++j) entropy += 0.5 + log_sqrt_2pi + log((double)scale[j]);
  const double expected = -(total / S + entropy);
  const double err = fabs((double)out[0] - expected) / fabs(expected);
  double num = 0.0, den = 0.0;
  for (int j = 0; j < P; ++j) {
    const double g = -grad_loc[j] / S;            /* d loss / d loc_j */
    num += (out[1 + j] - g) * (out[1 + j] - g);
    den += g * g;
  }
  printf("loss %.6f expected %.6f rel err %.2e | grad loc rel-l2 %.2e | status %u | %d kernels\n", out[0], expected, err,
         sqrt(num / den), status, launches);
  if (!(err < 1e-4) || !(sqrt(num / den) < 5e-3) || status != 0 || launches != 4) { printf("FAILED\n"); return 1; }

  /* the fused SVI step: transforms + evaluation + Adam, in place on `raw` */
  mnf_adam_t adam = {0.05f, 0.9f, 0.999f, 1e-8f, draw, dcode, dm, dv, dP, dsteps};
  float first = 0.f, last = 0.f;
  for (int step = 0; step < 60; ++step) {
    CHECK_MNF(mnf_svi_step(plan, &buffers, &adam, 1234u, 100u + step, MNF_STEP_ENTROPY, NULL));
    if (step == 0 || step == 59) {
      CHECK_CUDA(cudaDeviceSynchronize());
      CHECK_CUDA(cudaMemcpy(step == 0 ? &first : &last, dout, sizeof(float), cudaMemcpyDeviceToHost));
    }
  }
  int64_t steps = 0;
  CHECK_CUDA(cudaMemcpy(&steps, dsteps, sizeof(steps), cudaMemcpyDeviceToHost));
  CHECK_CUDA(cudaMemcpy(&status, dstatus, sizeof(status), cudaMemcpyDeviceToHost));
  printf("fused SVI: loss %.3f -> %.3f after %lld Adam steps, status %u\n", first, last, (long long)steps, status);
  if (!(last < first) || steps != 60 || status != 0) { printf("FAILED\n"); return 1; }
  CHECK_MNF(mnf_plan_destroy(plan));
  printf("OK\n");
  return 0;
}
