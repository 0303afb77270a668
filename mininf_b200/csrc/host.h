// Host-side helpers shared by the translation units of the library (abi.cu, dense.cu, site.cu,
// rowlatent.cu): the per-device property cache and the partial-sum reduction launcher.
#pragma once

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace mnf {

struct DeviceCache {
  bool ready = false;
  int sm_count = 0;
  int cc_major = 0;
  int cc_minor = 0;
  int max_smem_optin = 0;
  size_t total_mem = 0;
};

// defined in abi.cu
int device_cache(int device, DeviceCache** out);

inline int max_ctas(const DeviceCache& c) { return 4 * c.sm_count; }

// Column map of a sweep's per-CTA partial sums [n_cta][S][ncol] into the step accumulator:
// partial column 0 is the log-density; then `n_vec` consecutive latent columns starting at
// vec_lat (theta of a dense site), then up to 16 individually mapped scalar latent columns.
struct ColMap {
  int32_t n_vec;
  int32_t vec_lat;
  int32_t n_scalar;
  int32_t scalar_lat[16];
};

// fixed-order reduction of the partials into acc (small.cuh::reduce_partials_kernel; abi.cu)
int launch_reduce(const float* partial, int n_cta, int S, int ncol, const ColMap& map, double weight,
                  int D, double* acc, cudaStream_t stream);

// row-latent sweep with an optional fused optimiser (rowlatent.cu; RowAdam in common.cuh)
int rowlatent_sweep(const mnf_rowlatent_t* desc, const float* z, int n_particles, int n_latent_total, uint64_t seed,
                    uint64_t offset, const uint64_t* offset_dev, int with_entropy, double* acc, void* workspace,
                    size_t workspace_bytes, uint32_t* status, void* stream, const RowAdam& adam);

}  // namespace mnf
