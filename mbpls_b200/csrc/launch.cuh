// Host-side launch helpers shared by the .cu files (device properties, slab-shape choice).
#pragma once
#include <stdlib.h>
#include "stream.cuh"

namespace mbpls {

inline int smem_optin() {
  static int v = -1;
  if (v < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  return v;
}
inline int num_sms() {
  static int v = -1;
  if (v < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
  }
  return v;
}

// Bytes per bulk-copy op (multiple of 16).  One 80 KB op streams at only ~13 GB/s per SM (measured,
// profiles/), several smaller ops in flight are needed to reach the HBM rate.  MBPLS_BULK_CHUNK overrides.
inline int bulk_chunk_bytes() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MBPLS_BULK_CHUNK");
    long c = e ? atol(e) : 8192;
    if (c < 512) c = 512;
    v = static_cast<int>((c / 16) * 16);
  }
  return v;
}

// Choose the slab shape for the feature-resident pipeline.  Returns false if one feature does not
// fit twice in shared memory (-> caller uses the global-memory fallback).
inline bool pick_stream_shape(long ld, int p, StreamShape* sh, bool* cta_wide) {
  const size_t feat = static_cast<size_t>(ld) * sizeof(double);
  const size_t cap = static_cast<size_t>(smem_optin()) - 2048;  // static smem + barriers
  if (2 * feat > cap) return false;
  int G = 1, stages = 2;
  if (feat <= 8192) {  // short features: many per slab, one warp per feature, ~32 KB slabs, 3 stages
    G = static_cast<int>(32768 / feat);
    if (G > 64) G = 64;
    if (G < 8) G = 8;
    stages = 3;
    *cta_wide = false;
  } else {  // long features: the whole CTA works on one resident feature
    G = 1;
    stages = static_cast<int>(cap / feat);
    if (stages > 3) stages = 3;
    *cta_wide = true;
  }
  if (G > p) G = p > 0 ? p : 1;
  sh->ld = ld;
  sh->p = p;
  sh->G = G;
  sh->stages = stages;
  sh->chunk = bulk_chunk_bytes();
  return true;
}

inline int stream_grid(const StreamShape& sh, size_t smem) {
  const int ngroups = (sh.p + sh.G - 1) / sh.G;
  int per_sm = static_cast<int>(static_cast<size_t>(smem_optin()) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  int grid = num_sms() * per_sm;
  if (grid > ngroups) grid = ngroups;
  return grid < 1 ? 1 : grid;
}

}  // namespace mbpls
