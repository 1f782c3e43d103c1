// Feature-resident streaming pipeline: a persistent CTA pulls groups of G consecutive features
// (one contiguous slab of G*ld doubles of the feature-major matrix) into a ring of shared-memory
// stages with 1-D bulk (TMA) copies tracked by mbarriers, hands each resident slab to `op`, and --
// when WRITEBACK -- pushes the modified slab back with a bulk store.  This is what makes
// "column statistics + standardise" and "loadings + rank-1 deflation" exactly 1 read + 1 write of X
// (SURVEY.md 8d): the whole feature sits in shared memory between the reduction and the update.
#pragma once
#include "common.cuh"

namespace mbpls {

struct StreamShape {
  long ld;       // leading dimension of Xt (doubles)
  int p;         // features in this (local) matrix
  int G;         // features per slab
  int stages;    // ring depth (>= 2)
};

// smem: [stages][G*ld] doubles, then stages mbarriers (8 B each)
__host__ __device__ inline size_t stream_smem_bytes(const StreamShape& s) {
  return static_cast<size_t>(s.stages) * s.G * s.ld * sizeof(double) + static_cast<size_t>(s.stages) * 8 + 16;
}

template <bool WRITEBACK, class Op>
__device__ __forceinline__ void stream_feature_slabs(double* __restrict__ Xt, const StreamShape sh, Op& op) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* slab0 = reinterpret_cast<double*>(smem_raw);
  const size_t slab_elems = static_cast<size_t>(sh.G) * sh.ld;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(sh.stages) * slab_elems * sizeof(double));

  const int ngroups = (sh.p + sh.G - 1) / sh.G;
  const int first = blockIdx.x;
  const int step = gridDim.x;
  const int nmine = first < ngroups ? (ngroups - first + step - 1) / step : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < sh.stages; ++s) mbar_init(&full[s], 1);
    fence_barrier_init();
  }
  __syncthreads();

  auto group_feats = [&](int k) {
    const int g = first + k * step;
    const int f0 = g * sh.G;
    return min(sh.G, sh.p - f0);
  };
  auto issue_load = [&](int k) {
    const int g = first + k * step;
    const int nf = group_feats(k);
    const uint32_t bytes = static_cast<uint32_t>(static_cast<size_t>(nf) * sh.ld * sizeof(double));
    const int s = k % sh.stages;
    mbar_arrive_expect_tx(&full[s], bytes);
    bulk_g2s(slab0 + static_cast<size_t>(s) * slab_elems, Xt + static_cast<size_t>(g) * sh.G * sh.ld, bytes, &full[s]);
  };

  if (threadIdx.x == 0) {
    const int pre = min(sh.stages - 1, nmine);
    for (int k = 0; k < pre; ++k) issue_load(k);
  }

  for (int k = 0; k < nmine; ++k) {
    const int s = k % sh.stages;
    const uint32_t parity = static_cast<uint32_t>((k / sh.stages) & 1);
    if (threadIdx.x == 0) {
      const int kn = k + sh.stages - 1;  // refills the stage consumed in iteration k-1
      if (kn < nmine) {
        if (WRITEBACK) bulk_wait_read<0>();  // the store issued in iteration k-1 has finished reading smem
        issue_load(kn);
      }
    }
    mbar_wait(&full[s], parity);
    const int g = first + k * step;
    const int nf = group_feats(k);
    double* slab = slab0 + static_cast<size_t>(s) * slab_elems;
    op(slab, g * sh.G, nf);  // may contain __syncthreads(); must be called by all threads
    if (WRITEBACK) {
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the bulk-copy engine
      __syncthreads();
      if (threadIdx.x == 0) {
        bulk_s2g(Xt + static_cast<size_t>(g) * sh.G * sh.ld, slab,
                 static_cast<uint32_t>(static_cast<size_t>(nf) * sh.ld * sizeof(double)));
        bulk_commit();
      }
    } else {
      __syncthreads();  // all readers done before the stage is refilled
    }
  }
  if (WRITEBACK && threadIdx.x == 0) bulk_wait_all<0>();
}

}  // namespace mbpls
