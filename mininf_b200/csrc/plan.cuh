// One call per step: mnf_plan_create / mnf_elbo_fwd_bwd / mnf_svi_step / mnf_plan_destroy and the
// peer-memory exchange (mnf_xrank_*) of include/mininf_b200.h. The plan keeps host and device
// copies of the flat site tables the host traced once; a step call enqueues
//
//   rsample [+ parameter transforms]            small.cuh::rsample_kernel
//   sweeps over the observed sites              dense.cu / site.cu / rowlatent.cu (+ partial reductions)
//   short observed sites                        small.cuh::small_sites_kernel
//   [push acc to every peer rank]               small.cuh::xrank_push_kernel
//   tail: [gather peers] + prior sites + finalize [+ Adam]   small.cuh::tail_kernel (one block)
//
// on the caller's stream and returns; nothing synchronises, so the whole step is capturable in a
// CUDA graph. Replaces the per-step Python of EvidenceLowerBoundLoss.forward (mininf/nn.py:212-228)
// and, with mnf_svi_step, the zero_grad / backward / optimizer.step() around it (README.md:63-69).
// Included by abi.cu (one translation unit with the O(S*D) kernels of small.cuh).
#pragma once

#include <vector>

#include "host.h"
#include "small.cuh"

struct mnf_xrank {
  int world = 1, rank = 0;
  int64_t n = 0;                       // doubles per accumulator
  int device = 0;
  char* local = nullptr;               // cudaMalloc: [flags: world x u64, padded to 256 B][inbox: 2 x world x n doubles]
  size_t inbox_offset = 0;
  std::vector<char*> peers;            // base pointer of every rank's allocation (own: local)
  double** peer_inbox_dev = nullptr;   // device arrays [world]
  uint64_t** peer_flags_dev = nullptr;
  uint64_t* epoch_dev = nullptr;
  bool connected = false;
};

struct mnf_plan {
  int S = 0, D = 0, device = 0;
  uint32_t flags = 0;
  std::vector<mnf_latent_t> latents;
  std::vector<mnf_dense_site_t> dense;
  std::vector<int32_t> dense_mode;
  std::vector<mnf_site_t> group_sites;
  std::vector<int32_t> group_sizes;
  std::vector<mnf_site_t> small_observed, small_global;
  std::vector<mnf_rowlatent_t> rowlatent;
  // device copies of the tables the small kernels read
  mnf_latent_t* latents_dev = nullptr;
  mnf_site_t* small_observed_dev = nullptr;
  mnf_site_t* small_global_dev = nullptr;
  int64_t small_observed_longest = 0, small_global_work = 0;
  size_t workspace_bytes = 0;
  int last_launches = 0;
};

namespace {

template <typename T>
int upload(T** dev, const std::vector<T>& host, cudaStream_t stream, bool async) {
  if (host.empty()) return MNF_OK;
  const size_t bytes = host.size() * sizeof(T);
  if (*dev == nullptr) MNF_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(dev), bytes));
  if (async) MNF_CUDA_CHECK(cudaMemcpyAsync(*dev, host.data(), bytes, cudaMemcpyHostToDevice, stream));   // pageable source: staged before return
  else MNF_CUDA_CHECK(cudaMemcpy(*dev, host.data(), bytes, cudaMemcpyHostToDevice));
  return MNF_OK;
}

int check_desc(const mnf_plan_desc_t* d) {
  if (!d) return fail(MNF_E_INVALID, "plan descriptor is null%s%s");
  if (d->n_particles < 1 || d->n_latent_total < 1 || d->n_latents < 1 || !d->latents)
    return fail(MNF_E_INVALID, "a plan needs at least one particle and one latent site%s%s");
  if (d->n_dense < 0 || d->n_groups < 0 || d->n_small_observed < 0 || d->n_small_global < 0 || d->n_rowlatent < 0)
    return fail(MNF_E_INVALID, "negative table size in the plan descriptor%s%s");
  if ((d->n_dense && (!d->dense || !d->dense_mode)) || (d->n_groups && (!d->group_sites || !d->group_sizes)) ||
      (d->n_small_observed && !d->small_observed) || (d->n_small_global && !d->small_global) ||
      (d->n_rowlatent && !d->rowlatent))
    return fail(MNF_E_INVALID, "a table of the plan descriptor is null%s%s");
  for (int g = 0; g < d->n_groups; ++g)
    if (d->group_sizes[g] < 1 || d->group_sizes[g] > MNF_MAX_FUSED_SITES)
      return fail(MNF_E_INVALID, "a sweep group needs 1..MNF_MAX_FUSED_SITES sites%s%s");
  int covered = 0;
  for (int i = 0; i < d->n_latents; ++i) {
    const mnf_latent_t& L = d->latents[i];
    if (L.family < MNF_NORMAL || L.family > MNF_BETA || L.numel < 1 || L.offset != covered || !L.p0 || !L.p1)
      return fail(MNF_E_INVALID, "latent table: families Normal/Gamma/Beta, contiguous offsets, non-null parameters%s%s");
    covered += L.numel;
  }
  if (covered != d->n_latent_total) return fail(MNF_E_INVALID, "latent table does not cover n_latent_total columns%s%s");
  return MNF_OK;
}

void copy_tables(mnf_plan* p, const mnf_plan_desc_t* d) {
  p->latents.assign(d->latents, d->latents + d->n_latents);
  p->dense.assign(d->dense, d->dense + d->n_dense);
  p->dense_mode.assign(d->dense_mode, d->dense_mode + d->n_dense);
  int total = 0;
  for (int g = 0; g < d->n_groups; ++g) total += d->group_sizes[g];
  p->group_sites.assign(d->group_sites, d->group_sites + total);
  p->group_sizes.assign(d->group_sizes, d->group_sizes + d->n_groups);
  p->small_observed.assign(d->small_observed, d->small_observed + d->n_small_observed);
  p->small_global.assign(d->small_global, d->small_global + d->n_small_global);
  p->rowlatent.assign(d->rowlatent, d->rowlatent + d->n_rowlatent);
  p->small_observed_longest = 0;
  for (const mnf_site_t& s : p->small_observed) p->small_observed_longest = std::max(p->small_observed_longest, s.numel);
  p->small_global_work = 0;
  for (const mnf_site_t& s : p->small_global) p->small_global_work += s.numel;
}

XrankArgs xrank_args(const mnf_xrank* xr) {
  XrankArgs a;
  a.world = 1; a.rank = 0; a.n = 0;
  a.inbox = nullptr; a.flags = nullptr; a.peer_inbox = nullptr; a.peer_flags = nullptr; a.epoch = nullptr;
  if (xr != nullptr && xr->world > 1) {
    a.world = xr->world;
    a.rank = xr->rank;
    a.n = xr->n;
    a.inbox = reinterpret_cast<double*>(xr->local + xr->inbox_offset);
    a.flags = reinterpret_cast<uint64_t*>(xr->local);
    a.peer_inbox = xr->peer_inbox_dev;
    a.peer_flags = xr->peer_flags_dev;
    a.epoch = xr->epoch_dev;
  }
  return a;
}

// A prior-site table above this many (element, particle) evaluations runs as its own multi-block
// kernel in front of the single-block tail.
constexpr int64_t kTailGlobalWork = 1 << 17;

int run_step(mnf_plan* p, const mnf_buffers_t* b, const mnf_adam_t* adam, uint64_t seed, uint64_t offset,
             uint32_t flags, cudaStream_t stream) {
  if (!p || !b) return fail(MNF_E_INVALID, "step: plan or buffers are null%s%s");
  if (!b->z || !b->noise || !b->acc || !b->out || !b->status || (!b->workspace && p->workspace_bytes))
    return fail(MNF_E_INVALID, "step: a required buffer is null%s%s");
  if (b->workspace_bytes < p->workspace_bytes) return fail(MNF_E_INVALID, "step: workspace smaller than mnf_plan_workspace_bytes%s%s");
  if (!p->rowlatent.empty() && !b->rows) return fail(MNF_E_INVALID, "step: row-latent buffers are missing%s%s");
  if ((flags & MNF_STEP_ALL) == 0) flags |= MNF_STEP_ALL;
  const int S = p->S, D = p->D;
  const int with_entropy = (flags & MNF_STEP_ENTROPY) ? 1 : 0;
  const mnf_xrank* xr = b->xrank;
  if (xr != nullptr && xr->world > 1) {
    if (!xr->connected) return fail(MNF_E_INVALID, "step: mnf_xrank_connect has not been called%s%s");
    if (xr->n != (int64_t)S * (D + 1)) return fail(MNF_E_INVALID, "step: the exchange was created for another accumulator size%s%s");
  }
  const int launches_before = g_launches;

  if (flags & MNF_STEP_PRE) {
    if (adam != nullptr) {
      if (!adam->raw || !adam->transform || !adam->m || !adam->v || !adam->constrained || !adam->step)
        return fail(MNF_E_INVALID, "mnf_svi_step: a field of mnf_adam_t is null%s%s");
      const int64_t total = (int64_t)S * (D + 1);
      const int grid = (int)std::min<int64_t>((total + 255) / 256, 1024);
      rsample_kernel<<<grid, 256, 0, stream>>>(p->latents_dev, (int)p->latents.size(), S, D, b->noise_in, seed, offset,
                                               b->step_counter, b->z, b->noise, b->acc, b->status, adam->raw,
                                               adam->transform, adam->constrained);
      MNF_LAUNCH_CHECK();
    } else {
      if (int rc = mnf_rsample(p->latents_dev, (int)p->latents.size(), S, D, b->noise_in, seed, offset, b->step_counter,
                               b->z, b->noise, b->acc, b->status, stream))
        return rc;
    }
    for (size_t i = 0; i < p->dense.size(); ++i)
      if (int rc = mnf_dense_sweep(&p->dense[i], p->dense_mode[i], b->z, S, D, b->acc, b->workspace, b->workspace_bytes,
                                   b->status, stream))
        return rc;
    const mnf_site_t* group = p->group_sites.data();
    for (size_t g = 0; g < p->group_sizes.size(); ++g) {
      if (int rc = mnf_site_sweep(group, p->group_sizes[g], b->z, S, D, b->acc, b->workspace, b->workspace_bytes,
                                  p->flags & MNF_SWEEP_CLOSED_FORM, b->status, stream))
        return rc;
      group += p->group_sizes[g];
    }
    if (!p->small_observed.empty())
      if (int rc = mnf_small_sites(p->small_observed_dev, (int)p->small_observed.size(), p->small_observed_longest, b->z, S,
                                   D, b->acc, b->status, stream))
        return rc;
    for (size_t i = 0; i < p->rowlatent.size(); ++i) {
      mnf_rowlatent_t desc = p->rowlatent[i];
      desc.loc = b->rows[i].loc;
      desc.scale = b->rows[i].scale;
      desc.grad_loc = b->rows[i].grad_loc;
      desc.grad_scale = b->rows[i].grad_scale;
      desc.eps = b->rows[i].eps;
      RowAdam ra;
      std::memset(&ra, 0, sizeof(ra));
      if (adam != nullptr) {
        // the N*p variational parameters are trained inside the sweep, element by element
        const mnf_row_buffers_t& rb = b->rows[i];
        ra.enabled = 1;
        ra.lr = adam->lr; ra.beta1 = adam->beta1; ra.beta2 = adam->beta2; ra.eps = adam->eps;
        ra.step = adam->step;
        ra.loc_rw = rb.loc_rw; ra.raw_scale = rb.raw_scale;
        ra.m_loc = rb.m_loc; ra.v_loc = rb.v_loc; ra.m_scale = rb.m_scale; ra.v_scale = rb.v_scale;
        desc.loc = rb.loc_rw;
      }
      if (int rc = rowlatent_sweep(&desc, b->z, S, D, seed, offset, b->step_counter, with_entropy, b->acc,
                                   b->workspace, b->workspace_bytes, b->status, stream, ra))
        return rc;
    }
  }

  if (flags & MNF_STEP_POST) {
    const XrankArgs xa = xrank_args(xr);
    if (xa.world > 1) {
      xrank_push_kernel<<<xa.world, kXrankThreads, 0, stream>>>(xa, b->acc);
      MNF_LAUNCH_CHECK();
    }
    int n_global_tail = (int)p->small_global.size();
    if (n_global_tail > 0 && p->small_global_work * S > kTailGlobalWork) {
      // a large prior table: gather first (the priors are counted once, after the exchange), then
      // the multi-block kernel, then the tail without either
      if (xa.world > 1) return fail(MNF_E_UNSUPPORTED, "step: prior sites this large are not supported together with the peer exchange%s%s");
      int64_t longest = 0;
      for (const mnf_site_t& s : p->small_global) longest = std::max(longest, s.numel);
      if (int rc = mnf_small_sites(p->small_global_dev, n_global_tail, longest, b->z, S, D, b->acc, b->status, stream)) return rc;
      n_global_tail = 0;
    }
    AdamArgs ad;
    ad.raw = nullptr;
    if (adam != nullptr) {
      ad.lr = adam->lr; ad.beta1 = adam->beta1; ad.beta2 = adam->beta2; ad.eps = adam->eps;
      ad.raw = adam->raw; ad.transform = adam->transform; ad.m = adam->m; ad.v = adam->v; ad.step = adam->step;
    }
    const size_t stage = tail_stage_bytes(S, D);
    if (stage <= (size_t)200 * 1024) {
      if (stage > (size_t)48 * 1024)
        MNF_CUDA_CHECK(cudaFuncSetAttribute(tail_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage));
      tail_kernel<true><<<1, kTailThreads, stage, stream>>>(xa, p->small_global_dev, n_global_tail, p->latents_dev,
                                                            (int)p->latents.size(), S, D, b->z, b->noise, b->acc,
                                                            with_entropy, b->out, b->step_counter, b->status, ad);
    } else {
      tail_kernel<false><<<1, kTailThreads, 0, stream>>>(xa, p->small_global_dev, n_global_tail, p->latents_dev,
                                                         (int)p->latents.size(), S, D, b->z, b->noise, b->acc,
                                                         with_entropy, b->out, b->step_counter, b->status, ad);
    }
    MNF_LAUNCH_CHECK();
  }
  p->last_launches = g_launches - launches_before;
  return MNF_OK;
}

}  // namespace

extern "C" {

int mnf_plan_create(const mnf_plan_desc_t* desc, mnf_plan_t** out) {
  if (!out) return fail(MNF_E_INVALID, "mnf_plan_create: out is null%s%s");
  *out = nullptr;
  if (int rc = check_desc(desc)) return rc;
  int device = desc->device;
  if (device < 0) MNF_CUDA_CHECK(cudaGetDevice(&device));
  int current = 0;
  MNF_CUDA_CHECK(cudaGetDevice(&current));
  if (current != device) return fail(MNF_E_INVALID, "mnf_plan_create: make the plan's device current first%s%s");
  mnf_plan* p = new mnf_plan();
  p->S = desc->n_particles;
  p->D = desc->n_latent_total;
  p->device = device;
  p->flags = desc->flags;
  copy_tables(p, desc);
  int rc = upload(&p->latents_dev, p->latents, nullptr, false);
  if (!rc) rc = upload(&p->small_observed_dev, p->small_observed, nullptr, false);
  if (!rc) rc = upload(&p->small_global_dev, p->small_global, nullptr, false);
  if (rc) {
    mnf_plan_destroy(p);
    return rc;
  }
  int widest = p->D;
  for (const mnf_rowlatent_t& r : p->rowlatent) widest = std::max(widest, r.p + 3);
  p->workspace_bytes = std::max<size_t>(mnf_workspace_bytes(p->S, widest, device), (size_t)1 << 20);
  *out = p;
  return MNF_OK;
}

int mnf_plan_update(mnf_plan_t* p, const mnf_plan_desc_t* desc, void* stream) {
  if (!p) return fail(MNF_E_INVALID, "mnf_plan_update: plan is null%s%s");
  if (int rc = check_desc(desc)) return rc;
  int total = 0;
  for (int g = 0; g < desc->n_groups; ++g) total += desc->group_sizes[g];
  if (desc->n_particles != p->S || desc->n_latent_total != p->D || desc->n_latents != (int)p->latents.size() ||
      desc->n_dense != (int)p->dense.size() || desc->n_groups != (int)p->group_sizes.size() ||
      total != (int)p->group_sites.size() || desc->n_small_observed != (int)p->small_observed.size() ||
      desc->n_small_global != (int)p->small_global.size() || desc->n_rowlatent != (int)p->rowlatent.size())
    return fail(MNF_E_INVALID, "mnf_plan_update: the table shapes differ from the plan's; create a new plan%s%s");
  const std::vector<mnf_site_t> observed_before = p->small_observed, global_before = p->small_global;
  const std::vector<mnf_latent_t> latents_before = p->latents;
  p->flags = desc->flags;
  copy_tables(p, desc);
  auto differs = [](const void* a, const void* b, size_t bytes) { return bytes != 0 && std::memcmp(a, b, bytes) != 0; };
  cudaStream_t s = (cudaStream_t)stream;
  if (differs(latents_before.data(), p->latents.data(), p->latents.size() * sizeof(mnf_latent_t)))
    if (int rc = upload(&p->latents_dev, p->latents, s, true)) return rc;
  if (differs(observed_before.data(), p->small_observed.data(), p->small_observed.size() * sizeof(mnf_site_t)))
    if (int rc = upload(&p->small_observed_dev, p->small_observed, s, true)) return rc;
  if (differs(global_before.data(), p->small_global.data(), p->small_global.size() * sizeof(mnf_site_t)))
    if (int rc = upload(&p->small_global_dev, p->small_global, s, true)) return rc;
  return MNF_OK;
}

int mnf_plan_workspace_bytes(const mnf_plan_t* p, size_t* bytes) {
  if (!p || !bytes) return fail(MNF_E_INVALID, "mnf_plan_workspace_bytes: null argument%s%s");
  *bytes = p->workspace_bytes;
  return MNF_OK;
}

int mnf_plan_launches(const mnf_plan_t* p, int* kernels) {
  if (!p || !kernels) return fail(MNF_E_INVALID, "mnf_plan_launches: null argument%s%s");
  *kernels = p->last_launches;
  return MNF_OK;
}

int mnf_plan_destroy(mnf_plan_t* p) {
  if (!p) return MNF_OK;
  cudaFree(p->latents_dev);
  cudaFree(p->small_observed_dev);
  cudaFree(p->small_global_dev);
  delete p;
  return MNF_OK;
}

int mnf_elbo_fwd_bwd(mnf_plan_t* plan, const mnf_buffers_t* buffers, uint64_t seed, uint64_t offset, uint32_t flags,
                     void* stream) {
  return run_step(plan, buffers, nullptr, seed, offset, flags, (cudaStream_t)stream);
}

int mnf_svi_step(mnf_plan_t* plan, const mnf_buffers_t* buffers, const mnf_adam_t* adam, uint64_t seed, uint64_t offset,
                 uint32_t flags, void* stream) {
  if (!adam) return fail(MNF_E_INVALID, "mnf_svi_step: adam is null%s%s");
  return run_step(plan, buffers, adam, seed, offset, flags | MNF_STEP_ALL, (cudaStream_t)stream);
}

// ---- peer-memory exchange -----------------------------------------------------------------------
int mnf_xrank_create(int world, int rank, int64_t n_doubles, mnf_xrank_t** out, void* handle_out) {
  if (!out || !handle_out) return fail(MNF_E_INVALID, "mnf_xrank_create: null argument%s%s");
  *out = nullptr;
  if (world < 2 || world > 64 || rank < 0 || rank >= world || n_doubles < 1)
    return fail(MNF_E_INVALID, "mnf_xrank_create: 2..64 ranks, 0 <= rank < world, a positive accumulator size%s%s");
  static_assert(sizeof(cudaIpcMemHandle_t) == MNF_XRANK_HANDLE_BYTES, "CUDA IPC handle size");
  mnf_xrank* xr = new mnf_xrank();
  xr->world = world;
  xr->rank = rank;
  xr->n = n_doubles;
  if (cudaGetDevice(&xr->device) != cudaSuccess) { delete xr; return fail(MNF_E_CUDA, "cudaGetDevice failed%s%s"); }
  xr->inbox_offset = ((size_t)world * sizeof(uint64_t) + 255) / 256 * 256;
  const size_t bytes = xr->inbox_offset + (size_t)2 * world * n_doubles * sizeof(double);
  cudaError_t err = cudaMalloc(reinterpret_cast<void**>(&xr->local), bytes);
  if (err == cudaSuccess) err = cudaMemset(xr->local, 0, bytes);
  if (err == cudaSuccess) err = cudaMalloc(reinterpret_cast<void**>(&xr->epoch_dev), sizeof(uint64_t));
  if (err == cudaSuccess) err = cudaMemset(xr->epoch_dev, 0, sizeof(uint64_t));
  cudaIpcMemHandle_t handle;
  if (err == cudaSuccess) err = cudaIpcGetMemHandle(&handle, xr->local);
  if (err == cudaSuccess) err = cudaDeviceSynchronize();
  if (err != cudaSuccess) {
    mnf_xrank_destroy(xr);
    return fail(MNF_E_CUDA, "mnf_xrank_create: %s%s", cudaGetErrorString(err));
  }
  std::memcpy(handle_out, &handle, sizeof(handle));
  *out = xr;
  return MNF_OK;
}

int mnf_xrank_connect(mnf_xrank_t* xr, const void* handles) {
  if (!xr || !handles) return fail(MNF_E_INVALID, "mnf_xrank_connect: null argument%s%s");
  if (xr->connected) return fail(MNF_E_INVALID, "mnf_xrank_connect: already connected%s%s");
  xr->peers.assign(xr->world, nullptr);
  for (int r = 0; r < xr->world; ++r) {
    if (r == xr->rank) {
      xr->peers[r] = xr->local;
      continue;
    }
    cudaIpcMemHandle_t handle;
    std::memcpy(&handle, static_cast<const char*>(handles) + (size_t)r * MNF_XRANK_HANDLE_BYTES, sizeof(handle));
    void* ptr = nullptr;
    const cudaError_t err = cudaIpcOpenMemHandle(&ptr, handle, cudaIpcMemLazyEnablePeerAccess);
    if (err != cudaSuccess) return fail(MNF_E_CUDA, "mnf_xrank_connect: cudaIpcOpenMemHandle failed: %s%s", cudaGetErrorString(err));
    xr->peers[r] = static_cast<char*>(ptr);
  }
  std::vector<double*> inbox(xr->world);
  std::vector<uint64_t*> flags(xr->world);
  for (int r = 0; r < xr->world; ++r) {
    inbox[r] = reinterpret_cast<double*>(xr->peers[r] + xr->inbox_offset);
    flags[r] = reinterpret_cast<uint64_t*>(xr->peers[r]);
  }
  MNF_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&xr->peer_inbox_dev), xr->world * sizeof(double*)));
  MNF_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&xr->peer_flags_dev), xr->world * sizeof(uint64_t*)));
  MNF_CUDA_CHECK(cudaMemcpy(xr->peer_inbox_dev, inbox.data(), xr->world * sizeof(double*), cudaMemcpyHostToDevice));
  MNF_CUDA_CHECK(cudaMemcpy(xr->peer_flags_dev, flags.data(), xr->world * sizeof(uint64_t*), cudaMemcpyHostToDevice));
  xr->connected = true;
  return MNF_OK;
}

int mnf_xrank_destroy(mnf_xrank_t* xr) {
  if (!xr) return MNF_OK;
  for (int r = 0; r < (int)xr->peers.size(); ++r)
    if (r != xr->rank && xr->peers[r] != nullptr) cudaIpcCloseMemHandle(xr->peers[r]);
  cudaFree(xr->peer_inbox_dev);
  cudaFree(xr->peer_flags_dev);
  cudaFree(xr->epoch_dev);
  cudaFree(xr->local);
  delete xr;
  return MNF_OK;
}

}  // extern "C"
