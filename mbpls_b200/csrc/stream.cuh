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
  int chunk;     // bytes per bulk-copy op (a slab moves as ceil(bytes/chunk) ops so several are in flight)
};

// smem: [stages][G*ld] doubles, then 2*stages mbarriers (full, done; 8 B each)
__host__ __device__ inline size_t stream_smem_bytes(const StreamShape& s) {
  return static_cast<size_t>(s.stages) * s.G * s.ld * sizeof(double) + static_cast<size_t>(s.stages) * 16 + 16;
}

// Warp-specialised: the LAST warp of the CTA is the producer (one elected lane issues every bulk load
// and bulk store), all other warps are consumers.  full[s] completes when a slab has landed; done[s]
// collects one arrival per consumer warp once the slab has been processed (and, when WRITEBACK, its
// generic-proxy writes have been fenced for the async proxy).  Consumers therefore never wait for a
// store, and the producer refills a stage as soon as the store that drains it has read shared memory.
// `op(slab, first_feature, nfeat)` is called by consumer threads only; block-wide steps inside it must
// use the consumer barrier (block_sum_consumers).
template <bool WRITEBACK, class Op>
__device__ __forceinline__ void stream_feature_slabs(double* __restrict__ Xt, const StreamShape sh, Op& op) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* slab0 = reinterpret_cast<double*>(smem_raw);
  const size_t slab_elems = static_cast<size_t>(sh.G) * sh.ld;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + static_cast<size_t>(sh.stages) * slab_elems * sizeof(double));
  uint64_t* done = full + sh.stages;

  const int ngroups = (sh.p + sh.G - 1) / sh.G;
  const int first = blockIdx.x;
  const int step = gridDim.x;
  const int nmine = first < ngroups ? (ngroups - first + step - 1) / step : 0;
  const int n_cons_warps = (blockDim.x >> 5) - 1;
  const bool is_producer = (threadIdx.x >> 5) == n_cons_warps;

  if (threadIdx.x == 0) {
    for (int s = 0; s < sh.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&done[s], n_cons_warps);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (is_producer) {
    if ((threadIdx.x & 31) == 0) {
      auto issue_load = [&](int k) {
        const int g = first + k * step;
        const int nf = min(sh.G, sh.p - g * sh.G);
        const uint32_t bytes = static_cast<uint32_t>(static_cast<size_t>(nf) * sh.ld * sizeof(double));
        const int s = k % sh.stages;
        mbar_arrive_expect_tx(&full[s], bytes);
        unsigned char* dst = reinterpret_cast<unsigned char*>(slab0 + static_cast<size_t>(s) * slab_elems);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(Xt + static_cast<size_t>(g) * sh.G * sh.ld);
        for (uint32_t o = 0; o < bytes; o += sh.chunk) bulk_g2s(dst + o, src + o, min(static_cast<uint32_t>(sh.chunk), bytes - o), &full[s]);
      };
      const int pre = min(sh.stages, nmine);
      for (int k = 0; k < pre; ++k) issue_load(k);
      for (int k = 0; k < nmine; ++k) {
        const int s = k % sh.stages;
        const uint32_t parity = static_cast<uint32_t>((k / sh.stages) & 1);
        const int kn = k + sh.stages;
        if (!WRITEBACK && kn >= nmine) break;  // nothing left to issue
        mbar_wait(&done[s], parity);           // consumers are finished with stage s
        if (WRITEBACK) {
          const int g = first + k * step;
          const int nf = min(sh.G, sh.p - g * sh.G);
          const uint32_t bytes = static_cast<uint32_t>(static_cast<size_t>(nf) * sh.ld * sizeof(double));
          unsigned char* dst = reinterpret_cast<unsigned char*>(Xt + static_cast<size_t>(g) * sh.G * sh.ld);
          const unsigned char* src = reinterpret_cast<const unsigned char*>(slab0 + static_cast<size_t>(s) * slab_elems);
          for (uint32_t o = 0; o < bytes; o += sh.chunk) bulk_s2g(dst + o, src + o, min(static_cast<uint32_t>(sh.chunk), bytes - o));
          bulk_commit();
        }
        if (kn < nmine) {
          if (WRITEBACK) bulk_wait_read<0>();  // the store has drained the stage
          issue_load(kn);
        }
      }
      if (WRITEBACK) bulk_wait_all<0>();
    }
    return;
  }

  for (int k = 0; k < nmine; ++k) {
    const int s = k % sh.stages;
    const uint32_t parity = static_cast<uint32_t>((k / sh.stages) & 1);
    mbar_wait(&full[s], parity);
    const int g = first + k * step;
    const int nf = min(sh.G, sh.p - g * sh.G);
    op(slab0 + static_cast<size_t>(s) * slab_elems, g * sh.G, nf);
    if (WRITEBACK) fence_proxy_async_smem();  // this thread's smem writes -> visible to the bulk-copy engine
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&done[s]);
  }
}

}  // namespace mbpls
