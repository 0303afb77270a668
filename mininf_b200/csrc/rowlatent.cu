// Row-latent sweep (csrc/rowlatent.cuh) and the mnf_rowlatent_sweep entry point of
// include/mininf_b200.h.
#include "host.h"
#include "rowlatent.cuh"

using namespace mnf;

namespace {

bool host_link_has_latent(const mnf_link_t& L) { return L.a_lat >= 0 || L.b_lat >= 0; }

template <int SP, bool FULL>
int launch_rowlatent_inst(const mnf_rowlatent_t& d, const float* z, int S, int D, int s_begin,
                          int first_pass, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                          int with_entropy, float* partial, uint32_t* status, int grid, cudaStream_t stream,
                          const RowAdam& adam) {
  auto kernel = rowlatent_kernel<SP, FULL>;
  const size_t smem = rowlatent_smem_bytes<SP>();
  MNF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernel<<<grid, kRowThreads, smem, stream>>>(d, z, S, D, s_begin, first_pass, seed, offset, offset_dev,
                                               with_entropy, partial, status, adam);
  MNF_LAUNCH_CHECK();
  return MNF_OK;
}

// CTAs of the row-latent kernel resident per SM (the smaller of its two variants): the sweep
// launches exactly one wave, since the rows are walked with a grid stride and a partial second
// wave would run at a fraction of the occupancy.
template <int SP>
int rowlatent_resident(int* out) {
  const size_t smem = rowlatent_smem_bytes<SP>();
  int full = 1, partial = 1;
  MNF_CUDA_CHECK(cudaFuncSetAttribute(rowlatent_kernel<SP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MNF_CUDA_CHECK(cudaFuncSetAttribute(rowlatent_kernel<SP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MNF_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&full, rowlatent_kernel<SP, true>, kRowThreads, smem));
  MNF_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&partial, rowlatent_kernel<SP, false>, kRowThreads, smem));
  *out = std::max(1, std::min(full, partial));
  return MNF_OK;
}

// FULL: all SP particle slots of the launch are in use (no masking code in the kernel)
template <int SP>
int launch_rowlatent(const mnf_rowlatent_t& d, const float* z, int S, int D, int s_begin,
                     int first_pass, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                     int with_entropy, float* partial, uint32_t* status, int grid, cudaStream_t stream,
                     const RowAdam& adam) {
  if (S - s_begin >= SP)
    return launch_rowlatent_inst<SP, true>(d, z, S, D, s_begin, first_pass, seed, offset, offset_dev,
                                           with_entropy, partial, status, grid, stream, adam);
  return launch_rowlatent_inst<SP, false>(d, z, S, D, s_begin, first_pass, seed, offset, offset_dev,
                                          with_entropy, partial, status, grid, stream, adam);
}


}  // namespace

namespace mnf {

// the sweep with an optional fused optimiser (mnf_svi_step, plan.cuh); `adam.enabled == 0`: gradients only
int rowlatent_sweep(const mnf_rowlatent_t* desc, const float* z, int n_particles, int n_latent_total, uint64_t seed,
                    uint64_t offset, const uint64_t* offset_dev, int with_entropy, double* acc, void* workspace,
                    size_t workspace_bytes, uint32_t* status, void* stream_, const RowAdam& adam) {
  if (!desc || !z || !acc || !workspace || !status)
    return fail(MNF_E_INVALID, "mnf_rowlatent_sweep: null pointer%s%s");
  const mnf_rowlatent_t d = *desc;
  const int S = n_particles, D = n_latent_total;
  if (adam.enabled) {
    if (!adam.loc_rw || !adam.raw_scale || !adam.m_loc || !adam.v_loc || !adam.m_scale || !adam.v_scale || !adam.step)
      return fail(MNF_E_INVALID, "mnf_svi_step: a row latent's optimiser buffers are missing%s%s");
    if (S > 32 || d.eps != nullptr)
      return fail(MNF_E_UNSUPPORTED, "mnf_svi_step: row latents are trained in the sweep only for single-pass "
                                     "sweeps (at most 32 particles) with in-kernel draws%s%s");
  } else if (!d.scale || !d.grad_loc || !d.grad_scale) {
    return fail(MNF_E_INVALID, "mnf_rowlatent_sweep: bad descriptor%s%s");
  }
  if (!d.loc || d.n_rows < 0 || S <= 0)
    return fail(MNF_E_INVALID, "mnf_rowlatent_sweep: bad descriptor%s%s");
  if (d.p < 1 || d.p > 32)
    return fail(MNF_E_UNSUPPORTED, "mnf_rowlatent_sweep: 1..32 features per row in this build%s%s");
  if (host_link_has_latent(d.prior_loc) || d.prior_loc.x || d.prior_scale.x || d.prior_scale.b_lat >= 0 ||
      (d.feat && (host_link_has_latent(d.feat_scale) || d.feat_scale.x)))
    return fail(MNF_E_UNSUPPORTED,
                "mnf_rowlatent_sweep: prior location and feature scale must be constants, the prior "
                "scale a constant or scalar latent%s%s");
  if (d.resp) {
    if (d.resp_family != MNF_POISSON && d.resp_family != MNF_NORMAL && d.resp_family != MNF_BERNOULLI_LOGITS)
      return fail(MNF_E_UNSUPPORTED, "mnf_rowlatent_sweep: response family%s%s");
    if (d.beta_lat < 0 || d.beta_lat + d.p > D || d.icpt_lat >= D)
      return fail(MNF_E_INVALID, "mnf_rowlatent_sweep: latent columns out of range%s%s");
  }
  if (d.n_rows == 0) return MNF_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  DeviceCache* c;
  if (int rc = device_cache(-1, &c)) return rc;
  const int ncol = 1 + d.p + 5;
  int sp = S <= 4 ? 4 : (S <= 8 ? 8 : (S <= 16 ? 16 : 32));
  if (const char* force = std::getenv("MNF_ROWLATENT_SP")) {   // developer override: particles per pass
    const int v = std::atoi(force);
    if (v == 4 || v == 8 || v == 16 || v == 32) sp = v;
  }
  int resident = 1;
  if (int rc = sp == 4 ? rowlatent_resident<4>(&resident) : sp == 8 ? rowlatent_resident<8>(&resident)
               : sp == 16 ? rowlatent_resident<16>(&resident) : rowlatent_resident<32>(&resident))
    return rc;
  const int grid = (int)std::min<int64_t>((d.n_rows + kRowWarps - 1) / kRowWarps,
                                          std::min(max_ctas(*c), resident * c->sm_count));
  if ((size_t)grid * S * ncol * sizeof(float) > workspace_bytes)
    return fail(MNF_E_INVALID, "mnf_rowlatent_sweep: workspace too small%s%s");
  float* partial = static_cast<float*>(workspace);
  for (int s_begin = 0, pass = 0; s_begin < S; s_begin += sp, ++pass) {
    int rc;
    if (sp == 4) rc = launch_rowlatent<4>(d, z, S, D, s_begin, pass == 0, seed, offset, offset_dev, with_entropy, partial, status, grid, stream, adam);
    else if (sp == 8) rc = launch_rowlatent<8>(d, z, S, D, s_begin, pass == 0, seed, offset, offset_dev, with_entropy, partial, status, grid, stream, adam);
    else if (sp == 16) rc = launch_rowlatent<16>(d, z, S, D, s_begin, pass == 0, seed, offset, offset_dev, with_entropy, partial, status, grid, stream, adam);
    else rc = launch_rowlatent<32>(d, z, S, D, s_begin, pass == 0, seed, offset, offset_dev, with_entropy, partial, status, grid, stream, adam);
    if (rc) return rc;
  }
  // physical partial layout: 0 log-density, 1..p beta gradient (zeros without a response), then
  // intercept, prior location, prior scale, feature scale, response scale
  ColMap map;
  map.n_vec = d.p;
  map.vec_lat = d.resp ? d.beta_lat : 0;
  map.n_scalar = 5;
  for (int i = 0; i < 16; ++i) map.scalar_lat[i] = -1;
  map.scalar_lat[0] = d.resp ? d.icpt_lat : -1;
  map.scalar_lat[1] = -1;                     // prior location: constant
  map.scalar_lat[2] = d.prior_scale.a_lat;
  map.scalar_lat[3] = -1;                     // feature scale: constant
  map.scalar_lat[4] = (d.resp && d.resp_family == MNF_NORMAL) ? d.resp_scale.a_lat : -1;
  return launch_reduce(partial, grid, S, ncol, map, 1.0, D, acc, stream);
}

}  // namespace mnf

extern "C" {

int mnf_rowlatent_sweep(const mnf_rowlatent_t* desc, const float* z, int n_particles,
                        int n_latent_total, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                        int with_entropy, double* acc, void* workspace, size_t workspace_bytes,
                        uint32_t* status, void* stream_) {
  RowAdam none;
  std::memset(&none, 0, sizeof(none));
  return rowlatent_sweep(desc, z, n_particles, n_latent_total, seed, offset, offset_dev, with_entropy, acc, workspace,
                         workspace_bytes, status, stream_, none);
}

}  // extern "C"
