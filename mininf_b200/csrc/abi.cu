// C-ABI entry points of the mininf_b200 engine (declared in include/mininf_b200.h): the O(S*D)
// kernels and shared host helpers. The sweeps live in dense.cu, site.cu and rowlatent.cu.
// Every entry validates its arguments, enqueues kernels on the caller's stream and returns;
// nothing here synchronises, allocates user-visible memory or throws.
#include "host.h"
#include "small.cuh"

using namespace mnf;

namespace {

DeviceCache g_dev[64];
std::mutex g_dev_mutex;

}  // namespace

namespace mnf {

int device_cache(int device, DeviceCache** out) {
  if (device < 0) MNF_CUDA_CHECK(cudaGetDevice(&device));
  if (device >= 64) return fail(MNF_E_INVALID, "device index out of range%s%s");
  std::lock_guard<std::mutex> lock(g_dev_mutex);
  DeviceCache& c = g_dev[device];
  if (!c.ready) {
    cudaDeviceProp prop;
    MNF_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    c.sm_count = prop.multiProcessorCount;
    c.cc_major = prop.major;
    c.cc_minor = prop.minor;
    c.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    c.total_mem = prop.totalGlobalMem;
    c.ready = true;
  }
  *out = &c;
  return MNF_OK;
}

int launch_reduce(const float* partial, int n_cta, int S, int ncol, const ColMap& map, double weight,
                  int D, double* acc, cudaStream_t stream) {
  const int total = S * ncol, per_block = kReduceThreads / 32;   // one warp per output
  reduce_partials_kernel<<<(total + per_block - 1) / per_block, kReduceThreads, 0, stream>>>(
      partial, n_cta, S, ncol, map, weight, D, acc);
  MNF_LAUNCH_CHECK();
  return MNF_OK;
}

}  // namespace mnf

extern "C" {

int mnf_abi_version(void) { return MNF_ABI_VERSION; }

const char* mnf_last_error(void) { return g_last_error; }

int mnf_device_info(int device, mnf_device_info_t* out) {
  if (out == nullptr) return fail(MNF_E_INVALID, "mnf_device_info: out is null%s%s");
  DeviceCache* c;
  if (int rc = device_cache(device, &c)) return rc;
  out->sm_count = c->sm_count;
  out->cc_major = c->cc_major;
  out->cc_minor = c->cc_minor;
  out->max_smem_optin = c->max_smem_optin;
  out->total_mem = (int64_t)c->total_mem;
  return MNF_OK;
}

size_t mnf_workspace_bytes(int n_particles, int n_latent_total, int device) {
  DeviceCache* c;
  if (device_cache(device, &c) != MNF_OK) return 0;
  const size_t ncol = (size_t)std::max(n_latent_total + 3, 1 + 4 * MNF_MAX_FUSED_SITES);
  // + the moment path's scratch (range partials, moment rows, flag word: < 256 KB) and one block of
  // statistics rows per Normal site (< 64 KB each), so both can be in flight at once (csrc/site.cu)
  return (size_t)max_ctas(*c) * (size_t)n_particles * ncol * sizeof(float) + ((size_t)1 << 20) +
         (size_t)c->sm_count * 4400 * sizeof(float);   // + per-CTA Gram statistics (csrc/dense_gram.cuh)
}

int mnf_rsample(const mnf_latent_t* latents_dev, int n_latents, int n_particles, int n_latent_total,
                const float* noise_in, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                float* z, float* noise_out, double* acc, uint32_t* status, void* stream) {
  if (!latents_dev || !z || !noise_out || !acc || !status || n_latents <= 0 || n_particles <= 0 ||
      n_latent_total <= 0)
    return fail(MNF_E_INVALID, "mnf_rsample: null pointer or empty latent table%s%s");
  const int64_t total = (int64_t)n_particles * (n_latent_total + 1);
  const int grid = (int)std::min<int64_t>((total + 255) / 256, 1024);
  rsample_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(latents_dev, n_latents, n_particles,
                                                         n_latent_total, noise_in, seed, offset,
                                                         offset_dev, z, noise_out, acc, status,
                                                         nullptr, nullptr, nullptr);
  MNF_LAUNCH_CHECK();
  return MNF_OK;
}

int mnf_small_sites(const mnf_site_t* sites_dev, int n_sites, int64_t max_numel, const float* z,
                    int n_particles, int n_latent_total, double* acc, uint32_t* status,
                    void* stream) {
  if (n_sites == 0) return MNF_OK;
  if (!sites_dev || !z || !acc || !status || n_sites < 0 || max_numel < 0)
    return fail(MNF_E_INVALID, "mnf_small_sites: null pointer or negative size%s%s");
  if (max_numel == 0) return MNF_OK;
  const int bx = (int)std::min<int64_t>((max_numel + kSmallThreads - 1) / kSmallThreads, 256);
  if (n_particles > 65535) return fail(MNF_E_UNSUPPORTED, "mnf_small_sites: too many particles%s%s");
  dim3 grid(bx, n_sites, n_particles);
  small_sites_kernel<<<grid, kSmallThreads, 0, (cudaStream_t)stream>>>(sites_dev, z, n_particles,
                                                                       n_latent_total, acc, status);
  MNF_LAUNCH_CHECK();
  return MNF_OK;
}

int mnf_finalize(const mnf_latent_t* latents_dev, int n_latents, int n_particles, int n_latent_total,
                 const float* z, const float* noise, const double* acc, int with_entropy, float* out,
                 uint64_t* step_counter, uint32_t* status, void* stream) {
  if (!latents_dev || !z || !noise || !acc || !out || !status || n_latents <= 0)
    return fail(MNF_E_INVALID, "mnf_finalize: null pointer or empty latent table%s%s");
  finalize_kernel<<<1, kFinalThreads, 0, (cudaStream_t)stream>>>(latents_dev, n_latents, n_particles,
                                                                 n_latent_total, z, noise, acc,
                                                                 with_entropy, out, step_counter, status);
  MNF_LAUNCH_CHECK();
  return MNF_OK;
}

int mnf_masked_count(const float* value, const uint8_t* mask, int64_t numel, int64_t* out,
                     void* stream) {
  if (!value || !out || numel < 0) return fail(MNF_E_INVALID, "mnf_masked_count: bad argument%s%s");
  MNF_CUDA_CHECK(cudaMemsetAsync(out, 0, 2 * sizeof(int64_t), (cudaStream_t)stream));
  if (numel == 0) return MNF_OK;
  const int grid = (int)std::min<int64_t>((numel + 255) / 256, 2048);
  masked_count_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      value, mask, numel, reinterpret_cast<unsigned long long*>(out));
  MNF_LAUNCH_CHECK();
  return MNF_OK;
}

}  // extern "C"

#include "plan.cuh"
#include "predict.cuh"

extern "C" int mnf_predictive(const mnf_pred_site_t* sites_dev, int n_sites, int n_samples, int n_columns, float* z,
                              uint64_t seed, uint64_t offset, uint32_t* status, void* stream) {
  if (!sites_dev || !z || !status || n_sites < 0 || n_samples < 0 || n_columns < 1)
    return mnf::fail(MNF_E_INVALID, "mnf_predictive: null pointer or bad size%s%s");
  if (n_sites == 0 || n_samples == 0) return MNF_OK;
  mnf::predictive_kernel<<<n_samples, mnf::kPredictThreads, 0, (cudaStream_t)stream>>>(sites_dev, n_sites, n_columns, z,
                                                                                     seed, offset, status);
  MNF_LAUNCH_CHECK();
  return MNF_OK;
}
