// Dense-link sweeps (csrc/dense_simt.cuh, dense_tc.cuh, dense_tcr.cuh): kernel selection, TMA
// tensor maps and the mnf_dense_sweep / mnf_dense_tf32_kernel entry points of include/mininf_b200.h.
#include "host.h"
#include "dense_simt.cuh"
#include "dense_tc.cuh"
#include "dense_tcr.cuh"
#include "dense_gram.cuh"
#include "dense_th.cuh"

using namespace mnf;

namespace {

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
int make_x_map(const mnf_dense_site_t& site, CUtensorMapSwizzle swizzle, CUtensorMap* map,
               CUtensorMapDataType type = CU_TENSOR_MAP_DATA_TYPE_TFLOAT32) {
  TensorMapEncodeFn encode;
  if (int rc = tensor_map_encoder(&encode)) return rc;
  const cuuint64_t dims[2] = {(cuuint64_t)site.p, (cuuint64_t)site.n_rows};
  const cuuint64_t strides[1] = {(cuuint64_t)site.ldx * sizeof(float)};
  const cuuint32_t box[2] = {32, (cuuint32_t)tc::kTileM};
  const cuuint32_t elem_strides[2] = {1, 1};
  const CUresult r = encode(map, type, 2, const_cast<float*>(site.X), dims,
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
  MNF_LAUNCH_CHECK();
  return MNF_OK;
}

// fp16-operand kernel (dense_th.cuh): one fp32 staging image per tile, converted in shared memory
template <int FAMILY, bool ICPT>
int launch_dense_th(const mnf_dense_site_t& site, const float* z, int S, int D, float* partial,
                    uint32_t* status, int grid, cudaStream_t stream) {
  CUtensorMap map_x;
  if (int rc = make_x_map(site, CU_TENSOR_MAP_SWIZZLE_128B, &map_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT32)) return rc;
  auto kernel = th::dense_th_kernel<FAMILY, ICPT>;
  MNF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)th::kSmemBytes));
  kernel<<<grid, th::kThreads, th::kSmemBytes, stream>>>(map_x, site, z, S, D, partial, status);
  MNF_LAUNCH_CHECK();
  return MNF_OK;
}

// Normal family, p <= 64: data-only Gram statistics (dense_gram.cuh), taken only in mode
// MNF_DENSE_TF32_CLOSED_FORM. Workspace: per-CTA statistics, their fp64 totals, then one row block
// [S][1 + p + 2] for the common reduction. MNF_DENSE_NO_GRAM=1 switches it off for every mode
// (developer A/B switch).
bool gram_disabled() {
  const char* v = std::getenv("MNF_DENSE_NO_GRAM");
  return v != nullptr && v[0] != '\0' && v[0] != '0';
}

// The Gram statistics do not depend on the particles: any S, with or without a row mask.
bool gram_shape(int family, int p) {
  return family == MNF_NORMAL && p > 0 && p <= tc::kP && p % 4 == 0 && !gram_disabled();
}

size_t gram_workspace_bytes(int grid, int S) {
  return ((size_t)grid * gram::kCtaFloats * sizeof(float) + 255) / 256 * 256 +
         ((size_t)gram::kCtaFloats * sizeof(double) + 255) / 256 * 256 + (size_t)S * (1 + tc::kP + 2) * sizeof(float);
}

int launch_dense_gram(const mnf_dense_site_t& site, const float* z, int S, int D, void* workspace, float** rows_out,
                      uint32_t* status, int grid, cudaStream_t stream) {
  CUtensorMap map_mn;
  if (int rc = make_x_map(site, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, &map_mn)) return rc;
  char* ws = static_cast<char*>(workspace);
  float* cta_out = reinterpret_cast<float*>(ws);
  const size_t off_total = ((size_t)grid * gram::kCtaFloats * sizeof(float) + 255) / 256 * 256;
  double* total = reinterpret_cast<double*>(ws + off_total);
  float* rows = reinterpret_cast<float*>(ws + off_total + ((size_t)gram::kCtaFloats * sizeof(double) + 255) / 256 * 256);
  const char* dev = std::getenv("MNF_GRAM_DEV_SKIP");   // timing experiments only (dense_gram.cuh)
  const uint32_t dev_skip = dev != nullptr ? (uint32_t)std::atoi(dev) : 0u;
  auto kernel = site.mask != nullptr ? gram::dense_gram_kernel<true> : gram::dense_gram_kernel<false>;
  MNF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gram::kSmemBytes));
  kernel<<<grid, gram::kThreads, gram::kSmemBytes, stream>>>(map_mn, site, z, S, D, cta_out, status, dev_skip);
  MNF_LAUNCH_CHECK();
  gram::gram_reduce_kernel<<<(gram::kCtaFloats + 255) / 256, 256, 0, stream>>>(cta_out, grid, total);
  MNF_LAUNCH_CHECK();
  gram::gram_finish_kernel<<<S, tc::kP, 0, stream>>>(site, total, z, S, D, rows, status);
  MNF_LAUNCH_CHECK();
  *rows_out = rows;
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
  sh.C = (p + tcr::kChunk - 1) / tcr::kChunk;            // the last chunk is zero-padded by TMA
  // tensor memory: two eta buffers of 2NS (hi, lo) columns + 2C gradient tiles of NS columns. Very
  // wide matrices (C = 7) only fit 16 particle slots: they run in passes of 16 particles.
  int per_pass = 32;
  if ((4 + 2 * sh.C) * 32 > (int)tcr::kTmemCols) per_pass = 16;
  sh.passes = (S + per_pass - 1) / per_pass;
  sh.s_pass = (S + sh.passes - 1) / sh.passes;
  sh.NS = sh.s_pass <= 16 ? 16 : 32;
  if ((4 + 2 * sh.C) * sh.NS > (int)tcr::kTmemCols) return false;
  // split what is left of shared memory between the two operand rings, K ring first
  if (const char* force = std::getenv("MNF_TCR_STAGES")) {   // developer override "k,mn" (timing experiments)
    int k = 0, mn = 0;
    if (std::sscanf(force, "%d,%d", &k, &mn) == 2 && k >= 2 && mn >= 2 && k <= tcr::kMaxStages && mn <= tcr::kMaxStages) {
      sh.k_stages = k;
      sh.mn_stages = mn;
      sh.smem = tcr::make_layout(sh.NS, sh.C, k, mn).total;
      if (sh.smem <= (size_t)max_smem_optin) {
        *out = sh;
        return true;
      }
    }
  }
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
  MNF_LAUNCH_CHECK();
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

}  // namespace

extern "C" {

int mnf_dense_tf32_kernel(int family, int p, int n_particles) {
  TcrShape sh;
  const int which = dense_tf32_kernel(family, p, n_particles, 232448 /* sm_100 opt-in shared memory */, &sh);
  return which != 0 ? which : (n_particles > 0 && gram_shape(family, p) ? 3 : 0);
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

  if (mode == MNF_DENSE_F16) {
    const bool aligned = (reinterpret_cast<uintptr_t>(s.X) % 16 == 0) && (s.ldx % 4 == 0) &&
                         s.n_rows < (int64_t)1 << 31;
    if (!aligned || c->cc_major != 10 || p != th::kP || S > th::kNS)
      return fail(MNF_E_UNSUPPORTED, "mnf_dense_sweep: F16 mode needs p == 64, S <= 64, 16-byte aligned rows and an "
                                     "sm_100 device%s%s");
    const bool has_icpt = s.icpt_lat >= 0 || s.icpt_const != 0.0f;
    const int64_t n_tiles = (s.n_rows + th::kTileM - 1) / th::kTileM;
    grid = (int)std::min<int64_t>(n_tiles, c->sm_count);
    if ((size_t)grid * S * ncol * sizeof(float) > workspace_bytes)
      return fail(MNF_E_INVALID, "mnf_dense_sweep: workspace too small%s%s");
    int rc;
    if (s.family == MNF_NORMAL)
      rc = has_icpt ? launch_dense_th<MNF_NORMAL, true>(s, z, S, D, partial, status, grid, stream)
                    : launch_dense_th<MNF_NORMAL, false>(s, z, S, D, partial, status, grid, stream);
    else if (s.family == MNF_BERNOULLI_LOGITS)
      rc = has_icpt ? launch_dense_th<MNF_BERNOULLI_LOGITS, true>(s, z, S, D, partial, status, grid, stream)
                    : launch_dense_th<MNF_BERNOULLI_LOGITS, false>(s, z, S, D, partial, status, grid, stream);
    else
      rc = has_icpt ? launch_dense_th<MNF_POISSON, true>(s, z, S, D, partial, status, grid, stream)
                    : launch_dense_th<MNF_POISSON, false>(s, z, S, D, partial, status, grid, stream);
    if (rc) return rc;
  } else if (mode == MNF_DENSE_TF32 || mode == MNF_DENSE_TF32_CLOSED_FORM) {
    const bool closed_form = mode == MNF_DENSE_TF32_CLOSED_FORM && gram_shape(s.family, p);
    const bool aligned = (reinterpret_cast<uintptr_t>(s.X) % 16 == 0) && (s.ldx % 4 == 0) &&
                         s.n_rows < (int64_t)1 << 31;
    const bool has_icpt = s.icpt_lat >= 0 || s.icpt_const != 0.0f;
    TcrShape sh;
    const int which = dense_tf32_kernel(s.family, p, S, c->max_smem_optin, &sh);
    const bool c2_shape = which == 1, wide_shape = which == 2;
    if (!aligned || c->cc_major != 10 || !(c2_shape || wide_shape || closed_form))
      return fail(MNF_E_UNSUPPORTED,
                  "mnf_dense_sweep: TF32 mode needs p == 64 with S <= 64, or p a multiple of 4 with "
                  "S <= 128 (passes of <= 32 particles, (4 + ceil(p/64) * 2) * 32 within 512 TMEM columns), "
                  "16-byte aligned rows and an sm_100 device%s%s");
    const int64_t n_tiles = (s.n_rows + tc::kTileM - 1) / tc::kTileM;
    grid = (int)std::min<int64_t>(n_tiles, c->sm_count);
    if ((size_t)grid * S * ncol * sizeof(float) > workspace_bytes)
      return fail(MNF_E_INVALID, "mnf_dense_sweep: workspace too small%s%s");
    int rc;
    if (closed_form && gram_workspace_bytes(grid, S) <= workspace_bytes) {
      float* rows = nullptr;
      if (int rg = launch_dense_gram(s, z, S, D, workspace, &rows, status, grid, stream)) return rg;
      return launch_reduce(rows, 1, S, ncol, map, s.weight, D, acc, stream);
    }
    if (!(c2_shape || wide_shape)) return fail(MNF_E_INVALID, "mnf_dense_sweep: workspace too small%s%s");
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
    MNF_LAUNCH_CHECK();
  } else {
    return fail(MNF_E_INVALID, "mnf_dense_sweep: unknown mode%s%s");
  }

  return launch_reduce(partial, grid, S, ncol, map, s.weight, D, acc, stream);
}

#ifdef MNF_TC_DEBUG
int mnf_debug_buffer(void* ptr) {
  float* p = static_cast<float*>(ptr);
  MNF_CUDA_CHECK(cudaMemcpyToSymbol(tc::g_tc_debug, &p, sizeof(p)));
  return MNF_OK;
}
#endif

}  // extern "C"
