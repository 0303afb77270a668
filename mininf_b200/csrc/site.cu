// Element-wise site sweeps (csrc/site_sweep.cuh) and the mnf_site_sweep entry point of
// include/mininf_b200.h.
#include <cstdlib>
#include <mutex>

#include "host.h"
#include "site_sweep.cuh"

using namespace mnf;

namespace {

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
  MNF_LAUNCH_CHECK();
  return MNF_OK;
}


template <int Q>
int launch_poisson_exp_q(const mnf_site_t& site, const float* z, int S, int D, float* partial,
                         uint32_t* status, const uint32_t* need_exact, int grid, cudaStream_t stream) {
  auto kernel = poisson_exp_kernel<Q>;
  const size_t smem = poisson_exp_smem_bytes<Q>();
  MNF_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kernel<<<grid, kSweepThreads, smem, stream>>>(site, z, S, D, partial, status, need_exact);
  MNF_LAUNCH_CHECK();
  return MNF_OK;
}

int launch_poisson_exp(const mnf_site_t& site, const float* z, int S, int D, float* partial,
                       uint32_t* status, const uint32_t* need_exact, int grid, cudaStream_t stream) {
  if (S <= 32) return launch_poisson_exp_q<1>(site, z, S, D, partial, status, need_exact, grid, stream);
  if (S <= 64) return launch_poisson_exp_q<2>(site, z, S, D, partial, status, need_exact, grid, stream);
  return launch_poisson_exp_q<4>(site, z, S, D, partial, status, need_exact, grid, stream);
}

// Developer A/B switch: MNF_POISSON_EXACT=1 keeps the per-particle MUFU kernel for every call.
bool poisson_moments_disabled() {
  const char* v = std::getenv("MNF_POISSON_EXACT");     // read per call so a test can flip it
  return v != nullptr && v[0] != '\0' && v[0] != '0';
}

// Poisson(exp(A_s + B_s x)) through the data-only Chebyshev moments of site_sweep.cuh. `scratch`
// (after the exact kernel's partial rows) holds the range partials, the per-CTA moment rows and
// the need_exact word the finish kernel sets; returns that word's address, or NULL when this
// site's layout does not qualify (no covariate, strided or misaligned data, scratch too small).
// Range slots (site_sweep.cuh::RangeSlot): 64 per device, keyed by (covariate, mask, element count).
// The array is allocated and a new key registered only outside stream captures (cudaMalloc and the
// slot's reset are not capturable); a site met for the first time inside a capture runs uncached.
struct RangeSlots {
  static constexpr int kSlots = 64;
  RangeSlot* dev = nullptr;
  int used = 0;
  struct Key { const void* x; const void* mask; int64_t n; } keys[kSlots];
};
RangeSlots g_range_slots[16];
std::mutex g_range_mutex;

bool range_cache_disabled() {
  const char* v = std::getenv("MNF_POISSON_NO_RANGE_CACHE");   // developer A/B switch
  return v != nullptr && v[0] != '\0' && v[0] != '0';
}

RangeSlot* range_slot(const mnf_site_t& site, cudaStream_t stream) {
  if (range_cache_disabled()) return nullptr;
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= 16) return nullptr;
  std::lock_guard<std::mutex> lock(g_range_mutex);
  RangeSlots& rs = g_range_slots[device];
  for (int i = 0; i < rs.used; ++i)
    if (rs.keys[i].x == site.param[0].x && rs.keys[i].mask == site.mask && rs.keys[i].n == site.numel) return rs.dev + i;
  // the cache is an optimisation: any failure below means "run uncached" and must not leave an error
  // behind for the launch checks of the sweep
  auto uncached = [] { (void)cudaGetLastError(); return static_cast<RangeSlot*>(nullptr); };
  cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(stream, &capturing) != cudaSuccess) return uncached();
  if (capturing != cudaStreamCaptureStatusNone) return nullptr;
  if (rs.dev == nullptr) {
    if (cudaMalloc(&rs.dev, sizeof(RangeSlot) * RangeSlots::kSlots) != cudaSuccess) { rs.dev = nullptr; return uncached(); }
    if (cudaMemset(rs.dev, 0, sizeof(RangeSlot) * RangeSlots::kSlots) != cudaSuccess) return uncached();
  }
  if (rs.used == RangeSlots::kSlots) {      // full: start over (the slots revalidate themselves)
    if (cudaMemset(rs.dev, 0, sizeof(RangeSlot) * RangeSlots::kSlots) != cudaSuccess) return uncached();
    rs.used = 0;
  }
  if (cudaMemsetAsync(rs.dev + rs.used, 0, sizeof(RangeSlot), stream) != cudaSuccess) return uncached();
  const int i = rs.used++;
  rs.keys[i] = {site.param[0].x, site.mask, site.numel};
  return rs.dev + i;
}

int launch_poisson_moments(const mnf_site_t& site, const float* z, int S, int D, double* acc, char* scratch,
                           size_t scratch_bytes, uint32_t* status, int sm_count, cudaStream_t stream,
                           uint32_t** need_exact_out) {
  *need_exact_out = nullptr;
  const mnf_link_t& L0 = site.param[0];
  const bool vec = L0.x != nullptr && L0.x_stride == 1 && reinterpret_cast<uintptr_t>(L0.x) % 16 == 0 &&
                   reinterpret_cast<uintptr_t>(site.value) % 16 == 0 &&
                   (site.mask == nullptr || reinterpret_cast<uintptr_t>(site.mask) % 4 == 0);
  if (!vec || poisson_moments_disabled()) return MNF_OK;
  const int64_t groups = (site.numel + 4 * kChebThreads - 1) / (4 * kChebThreads);
  const int range_grid = (int)std::max<int64_t>(1, std::min<int64_t>(groups, 8 * (int64_t)sm_count));
  const int moment_grid = (int)std::max<int64_t>(1, std::min<int64_t>(groups, 3 * (int64_t)sm_count));
  const size_t rows_bytes = sizeof(double) * kChebCols * (size_t)moment_grid;
  const size_t range_bytes = (sizeof(float) * 2 * (size_t)range_grid + 15) / 16 * 16;
  const size_t tmax_bytes = (sizeof(float) * (size_t)moment_grid + 15) / 16 * 16;
  if (rows_bytes + range_bytes + tmax_bytes + 16 > scratch_bytes) return MNF_OK;
  double* rows = reinterpret_cast<double*>(scratch);
  float* range_partial = reinterpret_cast<float*>(scratch + rows_bytes);
  float* tmax_partial = reinterpret_cast<float*>(scratch + rows_bytes + range_bytes);
  uint32_t* need_exact = reinterpret_cast<uint32_t*>(scratch + rows_bytes + range_bytes + tmax_bytes);
  RangeSlot* slot = range_slot(site, stream);
  const size_t smem = poisson_moment_smem_bytes();
  auto moment_kernel = poisson_moment_kernel;
  MNF_CUDA_CHECK(cudaFuncSetAttribute(moment_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  poisson_range_kernel<<<range_grid, kChebThreads, 0, stream>>>(site, range_partial, slot);
  MNF_LAUNCH_CHECK();
  moment_kernel<<<moment_grid, kChebThreads, smem, stream>>>(site, range_partial, range_grid, slot, z, S, D, rows,
                                                             tmax_partial, status);
  MNF_LAUNCH_CHECK();
  poisson_moment_finish_kernel<<<1, kChebThreads, 0, stream>>>(site, rows, moment_grid, range_partial, range_grid, slot,
                                                              tmax_partial, z, S, D, acc, need_exact);
  MNF_LAUNCH_CHECK();
  *need_exact_out = need_exact;
  return MNF_OK;
}

// Normal site with an identity location link and an element-independent scale: one data-only pass
// for six sufficient statistics (on `stream`, which may be the side stream of overlap_streams),
// then - launch_normal_finish, always on the caller's stream - the per-particle closed forms
// straight into acc.
int normal_stats_grid(const mnf_site_t& site, bool* vec_out, int sm_count) {
  const mnf_link_t& L0 = site.param[0];
  const bool vec = reinterpret_cast<uintptr_t>(site.value) % 16 == 0 &&
                   (L0.x == nullptr || (L0.x_stride == 1 && reinterpret_cast<uintptr_t>(L0.x) % 16 == 0)) &&
                   (site.mask == nullptr || reinterpret_cast<uintptr_t>(site.mask) % 4 == 0);
  const int64_t per_thread = vec ? 4 : 1;
  const int64_t want = (site.numel + kStatThreads * per_thread - 1) / (kStatThreads * per_thread);
  *vec_out = vec;
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, 8 * (int64_t)sm_count));
}

int launch_normal_stats(const mnf_site_t& site, bool vec, int grid, double* cta_stats, uint32_t* status,
                        cudaStream_t stream) {
  if (vec) normal_stats_kernel<true><<<grid, kStatThreads, 0, stream>>>(site, cta_stats, status);
  else normal_stats_kernel<false><<<grid, kStatThreads, 0, stream>>>(site, cta_stats, status);
  MNF_LAUNCH_CHECK();
  return MNF_OK;
}

int launch_normal_finish(const mnf_site_t& site, const double* cta_stats, int grid, const float* z, int S, int D,
                         double* acc, uint32_t* status, cudaStream_t stream) {
  normal_stats_finish_kernel<<<1, 32 * kStatCols, 0, stream>>>(site, cta_stats, grid, z, S, D, acc, status);
  MNF_LAUNCH_CHECK();
  return MNF_OK;
}

// A second stream per device (plus fork / join events) so that the HBM-bound statistics pass of a
// Normal site runs beside the issue-bound moment pass of a Poisson site. The fork / join pattern
// is legal inside a stream capture, so a graphed step keeps the overlap.
struct OverlapStreams {
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};

int overlap_streams(OverlapStreams** out) {
  static std::mutex mutex;
  static OverlapStreams table[64];
  int device = 0;
  MNF_CUDA_CHECK(cudaGetDevice(&device));
  if (device < 0 || device >= 64) return fail(MNF_E_INVALID, "device index out of range%s%s");
  std::lock_guard<std::mutex> lock(mutex);
  OverlapStreams& o = table[device];
  if (o.side == nullptr) {
    MNF_CUDA_CHECK(cudaStreamCreateWithFlags(&o.side, cudaStreamNonBlocking));
    MNF_CUDA_CHECK(cudaEventCreateWithFlags(&o.fork, cudaEventDisableTiming));
    MNF_CUDA_CHECK(cudaEventCreateWithFlags(&o.join, cudaEventDisableTiming));
  }
  *out = &o;
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

}  // namespace

extern "C" {

int mnf_site_sweep(const mnf_site_t* sites, int n_sites, const float* z, int n_particles,
                   int n_latent_total, double* acc, void* workspace, size_t workspace_bytes,
                   uint32_t flags, uint32_t* status, void* stream_) {
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
  int n_generic = 0, n_normal = 0, n_poisson = 0;
  int normal_idx[MNF_MAX_FUSED_SITES], poisson_idx[MNF_MAX_FUSED_SITES];
  // Without MNF_SWEEP_CLOSED_FORM every site is evaluated once per (particle, element): Normal
  // sites stay in the fused per-particle kernel and the Poisson moment path is off.
  const bool closed_form = (flags & MNF_SWEEP_CLOSED_FORM) != 0;
  for (int i = 0; i < n_sites; ++i) {
    int kind = site_fast_kind(sites[i]);
    if (kind == kFastNormalId && !closed_form) kind = kFastNone;
    if (kind == kFastNone) generic[n_generic++] = sites[i];
    else if (kind == kFastNormalId) normal_idx[n_normal++] = i;
    else poisson_idx[n_poisson++] = i;
  }

  // Workspace plan: the Poisson kernels use the front (per-CTA partial rows of the per-particle
  // kernel, then the scratch of the moment path); the Normal statistics rows are carved from the
  // back so that both families can be in flight at once.
  const int pgrid = (int)std::min<int64_t>((n_chunks + kSweepWarps - 1) / kSweepWarps,
                                           (int64_t)pois_min_blocks(S <= 32 ? 1 : (S <= 64 ? 2 : 4)) * c->sm_count);
  const size_t exact_bytes = n_poisson ? ((size_t)pgrid * S * 5 * sizeof(float) + 255) / 256 * 256 : 0;
  if (exact_bytes > workspace_bytes) return fail(MNF_E_INVALID, "mnf_site_sweep: workspace too small%s%s");
  size_t back = workspace_bytes / 256 * 256;
  double* normal_rows[MNF_MAX_FUSED_SITES];
  int normal_grid[MNF_MAX_FUSED_SITES];
  bool normal_vec[MNF_MAX_FUSED_SITES];
  bool disjoint = true;                      // every Normal site got rows of its own behind the Poisson scratch
  for (int k = 0; k < n_normal; ++k) {
    normal_grid[k] = normal_stats_grid(sites[normal_idx[k]], &normal_vec[k], c->sm_count);
    const size_t bytes = ((size_t)normal_grid[k] * kStatCols * sizeof(double) + 255) / 256 * 256;
    if (back >= bytes && back - bytes >= exact_bytes + ((size_t)1 << 18)) {
      back -= bytes;
      normal_rows[k] = reinterpret_cast<double*>(static_cast<char*>(workspace) + back);
    } else {                                 // small workspace: one shared row block at the front, no overlap
      disjoint = false;
      const int64_t fits = (int64_t)(workspace_bytes / (kStatCols * sizeof(double)));
      if (fits < 1) return fail(MNF_E_INVALID, "mnf_site_sweep: workspace too small%s%s");
      normal_grid[k] = (int)std::min<int64_t>(normal_grid[k], fits);
      normal_rows[k] = static_cast<double*>(workspace);
    }
  }
  const bool overlap = disjoint && n_normal > 0 && n_poisson > 0;
  OverlapStreams* os = nullptr;
  if (overlap) {
    if (int rc = overlap_streams(&os)) return rc;
    MNF_CUDA_CHECK(cudaEventRecord(os->fork, stream));
    MNF_CUDA_CHECK(cudaStreamWaitEvent(os->side, os->fork, 0));
    for (int k = 0; k < n_normal; ++k)
      if (int rc = launch_normal_stats(sites[normal_idx[k]], normal_vec[k], normal_grid[k], normal_rows[k], status, os->side))
        return rc;
    MNF_CUDA_CHECK(cudaEventRecord(os->join, os->side));
  }
  for (int k = 0; k < n_poisson; ++k) {
    const mnf_site_t& site = sites[poisson_idx[k]];
    // data-only moment path first; the per-particle kernel runs only if its check fails (device flag)
    uint32_t* need_exact = nullptr;
    const size_t scratch_bytes = back - exact_bytes;
    if (closed_form)
      if (int rc = launch_poisson_moments(site, z, S, D, acc, static_cast<char*>(workspace) + exact_bytes,
                                          scratch_bytes, status, c->sm_count, stream, &need_exact))
        return rc;
    if (int rc = launch_poisson_exp(site, z, S, D, partial, status, need_exact, pgrid, stream)) return rc;
    ColMap fast_map;
    fast_map.n_vec = 0;
    fast_map.vec_lat = 0;
    fast_map.n_scalar = 4;
    for (int q = 0; q < 16; ++q) fast_map.scalar_lat[q] = -1;
    site_columns(site, fast_map.scalar_lat);
    if (int rr = launch_reduce(partial, pgrid, S, 5, fast_map, 1.0, D, acc, stream)) return rr;
  }
  if (overlap) MNF_CUDA_CHECK(cudaStreamWaitEvent(stream, os->join, 0));
  for (int k = 0; k < n_normal; ++k) {
    const mnf_site_t& site = sites[normal_idx[k]];
    if (!overlap)
      if (int rc = launch_normal_stats(site, normal_vec[k], normal_grid[k], normal_rows[k], status, stream)) return rc;
    if (int rc = launch_normal_finish(site, normal_rows[k], normal_grid[k], z, S, D, acc, status, stream)) return rc;
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

}  // extern "C"
