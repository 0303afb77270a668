// C-ABI entry points of the mininf_b200 engine (declared in include/mininf_b200.h).
// Every entry validates its arguments, enqueues kernels on the caller's stream and returns;
// nothing here synchronises, allocates user-visible memory or throws.
#include <algorithm>
#include <mutex>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "dense_simt.cuh"
#include "dense_tc.cuh"
#include "dense_tcr.cuh"
#include "rowlatent.cuh"
#include "site_sweep.cuh"
#include "small.cuh"

using namespace mnf;

namespace {

struct DeviceCache {
  bool ready = false;
  int sm_count = 0;
  int cc_major = 0;
  int cc_minor = 0;
  int max_smem_optin = 0;
  size_t total_mem = 0;
};

DeviceCache g_dev[64];
std::mutex g_dev_mutex;

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

inline int max_ctas(const DeviceCache& c) { return 4 * c.sm_count; }

int launch_reduce(const float* partial, int n_cta, int S, int ncol, const ColMap& map, double weight,
                  int D, double* acc, cudaStream_t stream) {
  const int total = S * ncol;
  reduce_partials_kernel<<<(total + 255) / 256, 256, 0, stream>>>(partial, n_cta, S, ncol, map,
                                                                  weight, D, acc);
  MNF_CUDA_CHECK(cudaGetLastError());
  return MNF_OK;
}

// cuTensorMapEncodeTiled is resolved through the runtime so the library has no link-time
// dependency on libcuda.
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                      const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                      CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int tensor_map_encoder(TensorMapEncodeFn* out) {
  static TensorMapEncodeFn fn = nullptr;
  static std::mutex m;
  std::lock_guard<std::mutex> lock(m);
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MNF_CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || ptr == nullptr)
      return fail(MNF_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver%s%s");
    fn = reinterpret_cast<TensorMapEncodeFn>(ptr);
  }
  *out = fn;
  return MNF_OK;
}

// X [n_rows][ldx] fp32 as a 2-D tensor (features fastest), boxes of 32 features x 128 rows,
// data type TFLOAT32: the TMA unit rounds to tf32 (nearest even) while copying.
int make_x_map(const mnf_dense_site_t& site, CUtensorMapSwizzle swizzle, CUtensorMap* map) {
  TensorMapEncodeFn encode;
  if (int rc = tensor_map_encoder(&encode)) return rc;
  const cuuint64_t dims[2] = {(cuuint64_t)site.p, (cuuint64_t)site.n_rows};
  const cuuint64_t strides[1] = {(cuuint64_t)site.ldx * sizeof(float)};
  const cuuint32_t box[2] = {32, (cuuint32_t)tc::kTileM};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_TFLOAT32, 2, const_cast<float*>(site.X), dims,
                            strides, box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MNF_E_CUDA, "cuTensorMapEncodeTiled failed%s%s");
  return MNF_OK;
}

template <int FAMILY, bool ICPT>
int launch_dense_tc(const mnf_dense_site_t& site, const float* z, int S, int D, float* partial,
                    uint32_t* status, int grid, cudaStream_t stream) {
  CUtensorMap map_k, map_mn;
  if (int rc = make_x_map(site, CU_TENSOR_MAP_SWIZZLE_128B, &map_k)) return rc;
  if (int rc = make_x_map(site, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, &map_mn)) return rc;
  auto kernel = tc::dense_tc_kernel<FAMILY, ICPT>;
  MNF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)tc::kSmemBytes));
  kernel<<<grid, tc::kThreads, tc::kSmemBytes, stream>>>(map_k, map_mn, site, z, S, D, partial, status);
  MNF_CUDA_CHECK(cudaGetLastError());
  return MNF_OK;
}

// rows-on-lanes tcgen05 kernel: p = 64 * C, S <= 32, optional intercept (dense_tcr.cuh)
struct TcrShape {
  int NS, C, k_stages, mn_stages;
  int passes, s_pass;     // more than 32 particles run as `passes` sweeps of at most s_pass particles
  size_t smem;
};
bool tcr_shape(int p, int S, int max_smem_optin, TcrShape* out) {
  if (p <= 0 || p % 4 != 0 || S > 128) return false;     // rows must be 16-byte multiples for TMA
  TcrShape sh;
  sh.passes = (S + 31) / 32;
  sh.s_pass = (S + sh.passes - 1) / sh.passes;
  sh.NS = sh.s_pass <= 16 ? 16 : 32;
  sh.C = (p + tcr::kChunk - 1) / tcr::kChunk;            // the last chunk is zero-padded by TMA
  if ((2 + 2 * sh.C) * sh.NS > (int)tcr::kTmemCols) return false;
  // split what is left of shared memory between the two operand rings, K ring first
  for (int stages = 2 * tcr::kMaxStages; stages >= 4; --stages) {
    sh.k_stages = (stages + 1) / 2;
    sh.mn_stages = stages / 2;
    sh.smem = tcr::make_layout(sh.NS, sh.C, sh.k_stages, sh.mn_stages).total;
    if (sh.smem <= (size_t)max_smem_optin) {
      *out = sh;
      return true;
    }
  }
  return false;
}

// 0 = no tcgen05 kernel for this shape, 1 = dense_tc.cuh, 2 = dense_tcr.cuh
int dense_tf32_kernel(int family, int p, int S, int max_smem_optin, TcrShape* sh) {
  if (family != MNF_NORMAL && family != MNF_BERNOULLI_LOGITS && family != MNF_POISSON) return 0;
  if (S <= 0) return 0;
  bool c2_shape = p == tc::kP && S <= tc::kNS;
  bool wide_shape = tcr_shape(p, S, max_smem_optin, sh);
  if (c2_shape && wide_shape) {
    // both cover p = 64 with S <= 32. Measured on B200: dense_tc.cuh wins for
    // the Normal and Poisson epilogues; the Bernoulli epilogue (exp, reciprocal, log per point) costs
    // per particle SLOT, and dense_tcr.cuh has 16 or 32 of them where dense_tc.cuh always has 64.
    const char* force = std::getenv("MNF_DENSE_TC_KERNEL");   // developer override: "tc" | "tcr"
    const bool prefer_wide = force ? std::strcmp(force, "tcr") == 0 : family == MNF_BERNOULLI_LOGITS;
    c2_shape = !prefer_wide;
    wide_shape = prefer_wide;
  }
  return c2_shape ? 1 : (wide_shape ? 2 : 0);
}

template <int FAMILY, int NS, bool ICPT>
int launch_dense_tcr_inst(const CUtensorMap& map_k, const CUtensorMap& map_mn, const mnf_dense_site_t& site,
                          const float* z, int S, int D, const TcrShape& sh, float* partial, uint32_t* status,
                          int grid, cudaStream_t stream) {
  auto kernel = tcr::dense_tcr_kernel<FAMILY, NS, ICPT>;
  MNF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh.smem));
  kernel<<<grid, tcr::kThreads, sh.smem, stream>>>(map_k, map_mn, site, z, S, D, sh.C, sh.k_stages, sh.mn_stages, partial, status);
  MNF_CUDA_CHECK(cudaGetLastError());
  return MNF_OK;
}

template <int FAMILY>
int launch_dense_tcr(const mnf_dense_site_t& site, const float* z, int S, int D, const TcrShape& sh,
                     bool has_icpt, float* partial, uint32_t* status, int grid, cudaStream_t stream) {
  CUtensorMap map_k, map_mn;
  if (int rc = make_x_map(site, CU_TENSOR_MAP_SWIZZLE_128B, &map_k)) return rc;
  if (int rc = make_x_map(site, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, &map_mn)) return rc;
  if (sh.NS == 16) {
    return has_icpt ? launch_dense_tcr_inst<FAMILY, 16, true>(map_k, map_mn, site, z, S, D, sh, partial, status, grid, stream)
                    : launch_dense_tcr_inst<FAMILY, 16, false>(map_k, map_mn, site, z, S, D, sh, partial, status, grid, stream);
  }
  return has_icpt ? launch_dense_tcr_inst<FAMILY, 32, true>(map_k, map_mn, site, z, S, D, sh, partial, status, grid, stream)
                  : launch_dense_tcr_inst<FAMILY, 32, false>(map_k, map_mn, site, z, S, D, sh, partial, status, grid, stream);
}

template <int NSITES>
int launch_site_sweep(const mnf_site_t* sites, const float* z, int S, int D, float* partial,
                      uint32_t* status, int grid, cudaStream_t stream) {
  SweepArgs<NSITES> args;
  for (int i = 0; i < NSITES; ++i) args.site[i] = sites[i];
  if (S <= 32) {
    auto kernel = site_sweep_kernel<NSITES, 1>;
    const size_t smem = site_sweep_smem_bytes<NSITES, 1>();
    MNF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, kSweepThreads, smem, stream>>>(args, z, S, D, partial, status);
  } else if (S <= 64) {
    auto kernel = site_sweep_kernel<NSITES, 2>;
    const size_t smem = site_sweep_smem_bytes<NSITES, 2>();
    MNF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, kSweepThreads, smem, stream>>>(args, z, S, D, partial, status);
  } else {
    auto kernel = site_sweep_kernel<NSITES, 4>;
    const size_t smem = site_sweep_smem_bytes<NSITES, 4>();
    MNF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<grid, kSweepThreads, smem, stream>>>(args, z, S, D, partial, status);
  }
  MNF_CUDA_CHECK(cudaGetLastError());
  return MNF_OK;
}


template <int Q>
int launch_poisson_exp_q(const mnf_site_t& site, const float* z, int S, int D, float* partial,
                         uint32_t* status, int grid, cudaStream_t stream) {
  auto kernel = poisson_exp_kernel<Q>;
  const size_t smem = poisson_exp_smem_bytes<Q>();
  MNF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernel<<<grid, kSweepThreads, smem, stream>>>(site, z, S, D, partial, status);
  MNF_CUDA_CHECK(cudaGetLastError());
  return MNF_OK;
}

int launch_poisson_exp(const mnf_site_t& site, const float* z, int S, int D, float* partial,
                       uint32_t* status, int grid, cudaStream_t stream) {
  if (S <= 32) return launch_poisson_exp_q<1>(site, z, S, D, partial, status, grid, stream);
  if (S <= 64) return launch_poisson_exp_q<2>(site, z, S, D, partial, status, grid, stream);
  return launch_poisson_exp_q<4>(site, z, S, D, partial, status, grid, stream);
}

// Normal site with an identity location link and an element-independent scale: one data-only pass
// for six sufficient statistics, then the per-particle closed forms straight into acc.
int launch_normal_stats(const mnf_site_t& site, const float* z, int S, int D, double* acc, void* workspace,
                        size_t workspace_bytes, uint32_t* status, int sm_count, cudaStream_t stream) {
  const mnf_link_t& L0 = site.param[0];
  const bool vec = reinterpret_cast<uintptr_t>(site.value) % 16 == 0 &&
                   (L0.x == nullptr || (L0.x_stride == 1 && reinterpret_cast<uintptr_t>(L0.x) % 16 == 0)) &&
                   (site.mask == nullptr || reinterpret_cast<uintptr_t>(site.mask) % 4 == 0);
  const int64_t per_thread = vec ? 4 : 1;
  const int64_t want = (site.numel + kStatThreads * per_thread - 1) / (kStatThreads * per_thread);
  const int64_t fits = (int64_t)(workspace_bytes / (kStatCols * sizeof(double)));   // one row of statistics per CTA
  if (fits < 1) return fail(MNF_E_INVALID, "mnf_site_sweep: workspace too small%s%s");
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(want, fits), 8 * (int64_t)sm_count));
  double* cta_stats = static_cast<double*>(workspace);
  if (vec) normal_stats_kernel<true><<<grid, kStatThreads, 0, stream>>>(site, cta_stats, status);
  else normal_stats_kernel<false><<<grid, kStatThreads, 0, stream>>>(site, cta_stats, status);
  MNF_CUDA_CHECK(cudaGetLastError());
  normal_stats_finish_kernel<<<1, 32 * kStatCols, 0, stream>>>(site, cta_stats, grid, z, S, D, acc, status);
  MNF_CUDA_CHECK(cudaGetLastError());
  return MNF_OK;
}

// scalar-latent columns of one site's four gradient sums (du0, du0*x0, du1, du1*x1)
void site_columns(const mnf_site_t& site, int32_t* cols) {
  const bool two = site.family <= MNF_BETA;
  cols[0] = site.param[0].a_lat;
  cols[1] = site.param[0].b_lat;
  cols[2] = two ? site.param[1].a_lat : -1;
  cols[3] = two ? site.param[1].b_lat : -1;
}

bool family_has_two_params(int family) { return family <= MNF_BETA; }

bool host_link_has_latent(const mnf_link_t& L) { return L.a_lat >= 0 || L.b_lat >= 0; }

template <int SP, bool FULL>
int launch_rowlatent_inst(const mnf_rowlatent_t& d, const float* z, int S, int D, int s_begin,
                          int first_pass, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                          int with_entropy, float* partial, uint32_t* status, int grid, cudaStream_t stream) {
  auto kernel = rowlatent_kernel<SP, FULL>;
  const size_t smem = rowlatent_smem_bytes<SP>();
  MNF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernel<<<grid, kRowThreads, smem, stream>>>(d, z, S, D, s_begin, first_pass, seed, offset, offset_dev,
                                               with_entropy, partial, status);
  MNF_CUDA_CHECK(cudaGetLastError());
  return MNF_OK;
}

// FULL: all SP particle slots of the launch are in use (no masking code in the kernel)
template <int SP>
int launch_rowlatent(const mnf_rowlatent_t& d, const float* z, int S, int D, int s_begin,
                     int first_pass, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                     int with_entropy, float* partial, uint32_t* status, int grid, cudaStream_t stream) {
  if (S - s_begin >= SP)
    return launch_rowlatent_inst<SP, true>(d, z, S, D, s_begin, first_pass, seed, offset, offset_dev,
                                           with_entropy, partial, status, grid, stream);
  return launch_rowlatent_inst<SP, false>(d, z, S, D, s_begin, first_pass, seed, offset, offset_dev,
                                          with_entropy, partial, status, grid, stream);
}


}  // namespace

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
  return (size_t)max_ctas(*c) * (size_t)n_particles * ncol * sizeof(float);
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
                                                         offset_dev, z, noise_out, acc, status);
  MNF_CUDA_CHECK(cudaGetLastError());
  return MNF_OK;
}

int mnf_dense_tf32_kernel(int family, int p, int n_particles) {
  TcrShape sh;
  return dense_tf32_kernel(family, p, n_particles, 232448 /* sm_100 opt-in shared memory */, &sh);
}

int mnf_dense_sweep(const mnf_dense_site_t* site, int mode, const float* z, int n_particles,
                    int n_latent_total, double* acc, void* workspace, size_t workspace_bytes,
                    uint32_t* status, void* stream_) {
  if (!site || !z || !acc || !workspace || !status)
    return fail(MNF_E_INVALID, "mnf_dense_sweep: null pointer%s%s");
  const mnf_dense_site_t s = *site;
  const int S = n_particles, D = n_latent_total, p = s.p;
  if (!s.X || !s.y || p <= 0 || s.n_rows < 0 || s.ldx < p || S <= 0)
    return fail(MNF_E_INVALID, "mnf_dense_sweep: bad site shape%s%s");
  if (s.family != MNF_NORMAL && s.family != MNF_BERNOULLI_LOGITS && s.family != MNF_POISSON)
    return fail(MNF_E_UNSUPPORTED, "mnf_dense_sweep: family has no dense-link kernel%s%s");
  if (s.theta_lat < 0 || s.theta_lat + p > D || s.icpt_lat >= D)
    return fail(MNF_E_INVALID, "mnf_dense_sweep: latent columns out of range%s%s");
  if (s.family == MNF_NORMAL && s.scale.x != nullptr)
    return fail(MNF_E_UNSUPPORTED, "mnf_dense_sweep: per-row scale is not supported%s%s");
  cudaStream_t stream = (cudaStream_t)stream_;
  DeviceCache* c;
  if (int rc = device_cache(-1, &c)) return rc;
  if (s.n_rows == 0) return MNF_OK;

  const int ncol = 1 + p + 2;
  float* partial = static_cast<float*>(workspace);
  int grid = 0;
  ColMap map;
  map.n_vec = p;
  map.vec_lat = s.theta_lat;
  map.n_scalar = 2;
  for (int i = 0; i < 16; ++i) map.scalar_lat[i] = -1;
  map.scalar_lat[0] = s.icpt_lat;
  // gradient w.r.t. the scale link's pre-transform value u goes to its latent scalar
  map.scalar_lat[1] = s.family == MNF_NORMAL ? s.scale.a_lat : -1;

  if (mode == MNF_DENSE_TF32) {
    const bool aligned = (reinterpret_cast<uintptr_t>(s.X) % 16 == 0) && (s.ldx % 4 == 0) &&
                         s.n_rows < (int64_t)1 << 31;
    const bool has_icpt = s.icpt_lat >= 0 || s.icpt_const != 0.0f;
    TcrShape sh;
    const int which = dense_tf32_kernel(s.family, p, S, c->max_smem_optin, &sh);
    const bool c2_shape = which == 1, wide_shape = which == 2;
    if (!aligned || c->cc_major != 10 || !(c2_shape || wide_shape))
      return fail(MNF_E_UNSUPPORTED,
                  "mnf_dense_sweep: TF32 mode needs p == 64 with S <= 64, or p a multiple of 4 with "
                  "S <= 128 (passes of <= 32 particles, (2 + ceil(p/64) * 2) * 32 within 512 TMEM columns), "
                  "16-byte aligned rows and an sm_100 device%s%s");
    const int64_t n_tiles = (s.n_rows + tc::kTileM - 1) / tc::kTileM;
    grid = (int)std::min<int64_t>(n_tiles, c->sm_count);
    if ((size_t)grid * S * ncol * sizeof(float) > workspace_bytes)
      return fail(MNF_E_INVALID, "mnf_dense_sweep: workspace too small%s%s");
    int rc;
    if (c2_shape) {
      if (s.family == MNF_NORMAL)
        rc = has_icpt ? launch_dense_tc<MNF_NORMAL, true>(s, z, S, D, partial, status, grid, stream)
                      : launch_dense_tc<MNF_NORMAL, false>(s, z, S, D, partial, status, grid, stream);
      else if (s.family == MNF_BERNOULLI_LOGITS)
        rc = has_icpt ? launch_dense_tc<MNF_BERNOULLI_LOGITS, true>(s, z, S, D, partial, status, grid, stream)
                      : launch_dense_tc<MNF_BERNOULLI_LOGITS, false>(s, z, S, D, partial, status, grid, stream);
      else
        rc = has_icpt ? launch_dense_tc<MNF_POISSON, true>(s, z, S, D, partial, status, grid, stream)
                      : launch_dense_tc<MNF_POISSON, false>(s, z, S, D, partial, status, grid, stream);
    } else {
      // at most 32 particles per sweep: larger S runs in passes (X is re-read by every pass)
      for (int pass = 0; pass < sh.passes; ++pass) {
        const int s0 = pass * sh.s_pass, sn = std::min(sh.s_pass, S - s0);
        const float* zp = z + (size_t)s0 * D;
        float* pp = partial + (size_t)grid * s0 * ncol;
        if (s.family == MNF_NORMAL) rc = launch_dense_tcr<MNF_NORMAL>(s, zp, sn, D, sh, has_icpt, pp, status, grid, stream);
        else if (s.family == MNF_BERNOULLI_LOGITS) rc = launch_dense_tcr<MNF_BERNOULLI_LOGITS>(s, zp, sn, D, sh, has_icpt, pp, status, grid, stream);
        else rc = launch_dense_tcr<MNF_POISSON>(s, zp, sn, D, sh, has_icpt, pp, status, grid, stream);
        if (rc) return rc;
        rc = launch_reduce(pp, grid, sn, ncol, map, s.weight, D, acc + (size_t)s0 * (D + 1), stream);
        if (rc) return rc;
      }
      return MNF_OK;
    }
    if (rc) return rc;
  } else if (mode == MNF_DENSE_FP32) {
    const size_t smem = dense_simt_smem_bytes(S, p);
    if (smem > (size_t)c->max_smem_optin)
      return fail(MNF_E_UNSUPPORTED, "mnf_dense_sweep: p x S too large for the fp32 kernel%s%s");
    const int64_t n_tiles = (s.n_rows + kSimtRows - 1) / kSimtRows;
    grid = (int)std::min<int64_t>(n_tiles, 2 * c->sm_count);
    if ((size_t)grid * S * ncol * sizeof(float) > workspace_bytes)
      return fail(MNF_E_INVALID, "mnf_dense_sweep: workspace too small%s%s");
    MNF_CUDA_CHECK(cudaFuncSetAttribute(dense_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dense_simt_kernel<<<grid, kSimtThreads, smem, stream>>>(s, z, S, D, partial, status);
    MNF_CUDA_CHECK(cudaGetLastError());
  } else {
    return fail(MNF_E_INVALID, "mnf_dense_sweep: unknown mode%s%s");
  }

  return launch_reduce(partial, grid, S, ncol, map, s.weight, D, acc, stream);
}

int mnf_site_sweep(const mnf_site_t* sites, int n_sites, const float* z, int n_particles,
                   int n_latent_total, double* acc, void* workspace, size_t workspace_bytes,
                   uint32_t* status, void* stream_) {
  if (!sites || !z || !acc || !workspace || !status)
    return fail(MNF_E_INVALID, "mnf_site_sweep: null pointer%s%s");
  if (n_sites < 1 || n_sites > MNF_MAX_FUSED_SITES)
    return fail(MNF_E_INVALID, "mnf_site_sweep: 1..MNF_MAX_FUSED_SITES sites per call%s%s");
  const int S = n_particles, D = n_latent_total;
  if (S <= 0 || S > 128) return fail(MNF_E_UNSUPPORTED, "mnf_site_sweep: 1..128 particles%s%s");
  for (int i = 0; i < n_sites; ++i) {
    const mnf_site_t& st = sites[i];
    if (st.numel != sites[0].numel || st.value == nullptr || st.value_lat >= 0)
      return fail(MNF_E_INVALID, "mnf_site_sweep: fused sites need observed values of equal length%s%s");
    if (st.family < 0 || st.family >= MNF_NUM_FAMILIES)
      return fail(MNF_E_INVALID, "mnf_site_sweep: unknown family%s%s");
    for (int p = 0; p < 2; ++p) {
      const mnf_link_t& L = st.param[p];
      if ((L.a_lat >= 0 && L.a_stride != 0) || (L.b_lat >= 0 && L.b_stride != 0))
        return fail(MNF_E_UNSUPPORTED, "mnf_site_sweep: links must reference scalar latents%s%s");
      if (L.a_lat >= D || L.b_lat >= D)
        return fail(MNF_E_INVALID, "mnf_site_sweep: latent column out of range%s%s");
    }
  }
  if (sites[0].numel == 0) return MNF_OK;
  cudaStream_t stream = (cudaStream_t)stream_;
  DeviceCache* c;
  if (int rc = device_cache(-1, &c)) return rc;
  const int64_t n_chunks = (sites[0].numel + 31) / 32;
  const int grid = (int)std::min<int64_t>((n_chunks + kSweepWarps - 1) / kSweepWarps, 2 * c->sm_count);
  float* partial = static_cast<float*>(workspace);

  // Sites with a specialised kernel (site_sweep.cuh: Poisson with an exp link, Normal with an
  // identity location link and a per-particle scale) run on their own; the rest stay fused.
  mnf_site_t generic[MNF_MAX_FUSED_SITES];
  int n_generic = 0;
  for (int i = 0; i < n_sites; ++i) {
    const int kind = site_fast_kind(sites[i]);
    if (kind == kFastNone) {
      generic[n_generic++] = sites[i];
      continue;
    }
    if (kind == kFastNormalId) {
      if (int rc = launch_normal_stats(sites[i], z, S, D, acc, workspace, workspace_bytes, status, c->sm_count, stream))
        return rc;
      continue;
    }
    const int pgrid = (int)std::min<int64_t>((n_chunks + kSweepWarps - 1) / kSweepWarps,
                                             (int64_t)pois_min_blocks(S <= 32 ? 1 : (S <= 64 ? 2 : 4)) * c->sm_count);
    if ((size_t)pgrid * S * 5 * sizeof(float) > workspace_bytes)
      return fail(MNF_E_INVALID, "mnf_site_sweep: workspace too small%s%s");
    if (int rc = launch_poisson_exp(sites[i], z, S, D, partial, status, pgrid, stream)) return rc;
    ColMap fast_map;
    fast_map.n_vec = 0;
    fast_map.vec_lat = 0;
    fast_map.n_scalar = 4;
    for (int k = 0; k < 16; ++k) fast_map.scalar_lat[k] = -1;
    site_columns(sites[i], fast_map.scalar_lat);
    if (int rr = launch_reduce(partial, pgrid, S, 5, fast_map, 1.0, D, acc, stream)) return rr;
  }
  if (n_generic == 0) return MNF_OK;
  sites = generic;
  n_sites = n_generic;
  const int n_templ = n_sites == 1 ? 1 : (n_sites == 2 ? 2 : 4);
  const int ncol = 1 + 4 * n_templ;
  if ((size_t)grid * S * ncol * sizeof(float) > workspace_bytes)
    return fail(MNF_E_INVALID, "mnf_site_sweep: workspace too small%s%s");

  // pad the site list to the template width with inert duplicates of zero weight
  mnf_site_t padded[MNF_MAX_FUSED_SITES];
  for (int i = 0; i < n_templ; ++i) {
    padded[i] = sites[i < n_sites ? i : 0];
    if (i >= n_sites) {
      padded[i].scale = 0.0;
      for (int p = 0; p < 2; ++p) { padded[i].param[p].a_lat = -1; padded[i].param[p].b_lat = -1; }
    }
  }
  int rc;
  if (n_templ == 1) rc = launch_site_sweep<1>(padded, z, S, D, partial, status, grid, stream);
  else if (n_templ == 2) rc = launch_site_sweep<2>(padded, z, S, D, partial, status, grid, stream);
  else rc = launch_site_sweep<4>(padded, z, S, D, partial, status, grid, stream);
  if (rc) return rc;

  ColMap map;
  map.n_vec = 0;
  map.vec_lat = 0;
  map.n_scalar = 4 * n_templ;
  for (int i = 0; i < 16; ++i) map.scalar_lat[i] = -1;
  for (int i = 0; i < n_sites; ++i) {
    const bool two = family_has_two_params(sites[i].family);
    map.scalar_lat[4 * i + 0] = sites[i].param[0].a_lat;
    map.scalar_lat[4 * i + 1] = sites[i].param[0].b_lat;
    map.scalar_lat[4 * i + 2] = two ? sites[i].param[1].a_lat : -1;
    map.scalar_lat[4 * i + 3] = two ? sites[i].param[1].b_lat : -1;
  }
  return launch_reduce(partial, grid, S, ncol, map, 1.0, D, acc, stream);
}

int mnf_rowlatent_sweep(const mnf_rowlatent_t* desc, const float* z, int n_particles,
                        int n_latent_total, uint64_t seed, uint64_t offset, const uint64_t* offset_dev,
                        int with_entropy, double* acc, void* workspace, size_t workspace_bytes,
                        uint32_t* status, void* stream_) {
  if (!desc || !z || !acc || !workspace || !status)
    return fail(MNF_E_INVALID, "mnf_rowlatent_sweep: null pointer%s%s");
  const mnf_rowlatent_t d = *desc;
  const int S = n_particles, D = n_latent_total;
  if (!d.loc || !d.scale || !d.grad_loc || !d.grad_scale || d.n_rows < 0 || S <= 0)
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
  const int grid = (int)std::min<int64_t>((d.n_rows + kRowWarps - 1) / kRowWarps, max_ctas(*c));
  if ((size_t)grid * S * ncol * sizeof(float) > workspace_bytes)
    return fail(MNF_E_INVALID, "mnf_rowlatent_sweep: workspace too small%s%s");
  float* partial = static_cast<float*>(workspace);
  int sp = S <= 4 ? 4 : (S <= 8 ? 8 : (S <= 16 ? 16 : 32));
  if (const char* force = std::getenv("MNF_ROWLATENT_SP")) {   // developer override: particles per pass
    const int v = std::atoi(force);
    if (v == 4 || v == 8 || v == 16 || v == 32) sp = v;
  }
  for (int s_begin = 0, pass = 0; s_begin < S; s_begin += sp, ++pass) {
    int rc;
    if (sp == 4) rc = launch_rowlatent<4>(d, z, S, D, s_begin, pass == 0, seed, offset, offset_dev, with_entropy, partial, status, grid, stream);
    else if (sp == 8) rc = launch_rowlatent<8>(d, z, S, D, s_begin, pass == 0, seed, offset, offset_dev, with_entropy, partial, status, grid, stream);
    else if (sp == 16) rc = launch_rowlatent<16>(d, z, S, D, s_begin, pass == 0, seed, offset, offset_dev, with_entropy, partial, status, grid, stream);
    else rc = launch_rowlatent<32>(d, z, S, D, s_begin, pass == 0, seed, offset, offset_dev, with_entropy, partial, status, grid, stream);
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
  MNF_CUDA_CHECK(cudaGetLastError());
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
  MNF_CUDA_CHECK(cudaGetLastError());
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
  MNF_CUDA_CHECK(cudaGetLastError());
  return MNF_OK;
}

#ifdef MNF_TC_DEBUG
int mnf_debug_buffer(void* ptr) {
  float* p = static_cast<float*>(ptr);
  MNF_CUDA_CHECK(cudaMemcpyToSymbol(tc::g_tc_debug, &p, sizeof(p)));
  return MNF_OK;
}
#endif

}  // extern "C"
