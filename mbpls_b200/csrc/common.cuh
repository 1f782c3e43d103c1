// Shared device helpers for the mbpls_b200 kernels (sm_100a only).
//
// Layout convention used by every kernel in this library ("feature-major"):
//   Xt  : p x ld doubles, one *feature* (a column of the reference's n x p matrix,
//         mbpls/mbpls.py:379) per row, ld >= n and ld % 16 == 0 so each feature starts on a
//         128-byte boundary and can be moved with one 1-D bulk (TMA) copy.
//   Yt  : q x ld, same convention.  n-vectors (u, ts, t_b) are plain arrays of length >= ld.
// Padding elements [n, ld) are zero and are never used in arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define MBPLS_FULL_MASK 0xffffffffu

namespace mbpls {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(MBPLS_FULL_MASK, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(MBPLS_FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(MBPLS_FULL_MASK, v, o));
  return v;
}
__device__ __forceinline__ int warp_or(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(MBPLS_FULL_MASK, v, o);
  return v;
}

// Deterministic block-wide sum of `NV` values per thread (blockDim <= 1024).  `scratch` must hold
// 32*NV doubles.  Every thread gets the result: warp shuffles, one smem exchange, and the same
// shuffle tree over the per-warp partials in every warp -> bitwise reproducible.  Two __syncthreads.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();  // protect scratch from a previous call's readers
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) scratch[k * 32 + warp] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(lane < nw ? scratch[k * 32 + lane] : 0.0);
}

// Same reduction among the consumer warps of a warp-specialised CTA (threads [0, nthreads), nthreads a
// multiple of 32): named barrier 1 instead of __syncthreads so the producer warp is not involved.
template <int NV>
__device__ __forceinline__ void block_sum_consumers(double (&v)[NV], double* scratch, int nthreads) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = nthreads >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) scratch[k * 32 + warp] = v[k];
  }
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(lane < nw ? scratch[k * 32 + lane] : 0.0);
}

// Reduction inside one 256-thread *group* of a CTA (group g uses named barrier 1+g and its own scratch):
// lets two independent feature pipelines share one CTA (and its shared-memory copies of ts / u0).
template <int NV>
__device__ __forceinline__ void group_sum256(double (&v)[NV], double* scratch, int group) {
  const int lane = threadIdx.x & 31, warp = (threadIdx.x & 255) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  asm volatile("bar.sync %0, 256;" ::"r"(group + 1) : "memory");
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) scratch[k * 8 + warp] = v[k];
  }
  asm volatile("bar.sync %0, 256;" ::"r"(group + 1) : "memory");
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double t = lane < 8 ? scratch[k * 8 + lane] : 0.0;
    t += __shfl_xor_sync(MBPLS_FULL_MASK, t, 4);
    t += __shfl_xor_sync(MBPLS_FULL_MASK, t, 2);
    t += __shfl_xor_sync(MBPLS_FULL_MASK, t, 1);
    v[k] = __shfl_sync(MBPLS_FULL_MASK, t, 0);
  }
}

__device__ __forceinline__ double block_sum1(double x, double* scratch) {
  double v[1] = {x};
  block_sum<1>(v, scratch);
  return v[0];
}

// ---- exact division by a divisor shared by many elements ---------------------------------------------------
// Standardisation divides a whole feature by its scale, the superlevel step divides n-vectors by their norms: one divisor,
// thousands of dividends.  An IEEE fp64 division is ~30 instructions on the FP64 pipe (MUFU.RCP64H seed, Newton steps,
// correction, range checks) and is what those passes stall on.  With r = RN(1/b) computed ONCE (a true division), Markstein's
// theorem gives the correctly rounded quotient in three instructions:
//     q0 = RN(a r);   e = a - b q0  (exact, one FMA);   q = RN(q0 + e r)  ==  RN(a / b)
// (q0 is within an ulp of a/b because r is correctly rounded; the FMA residual is exact; the final FMA rounds once.)  The
// result is bit-identical to `a / b`, so nothing downstream -- trip counts included -- can tell the difference.  Quotients that
// are zero, denormal-range, huge or infinite, and divisors far from 1 in magnitude, take the ordinary division (NaN stays NaN
// on the fast path: in NaN mode every warp holds missing values, and sending them to the slow path made every warp run both).
struct UniformDivisor {
  double b, r;
  bool fast;
  __device__ __forceinline__ explicit UniformDivisor(double divisor) : b(divisor), r(1.0 / divisor) {
    const double ab = fabs(divisor);
    fast = ab >= 0x1p-200 && ab <= 0x1p200;
  }
  __device__ __forceinline__ double operator()(double a) const {
    if (fast) {
      const double q0 = a * r;
      const double q = fma(fma(-q0, b, a), r, q0);
      const double aq = fabs(q);
      if (!(aq < 0x1p-700) && !(aq > 0x1p700)) return q;  // ordinary quotient -- or NaN (a missing value stays on the fast path)
    }
    return a / b;
  }
};

// ---- mbarrier + 1-D bulk async copies (TMA engine; SASS: UBLKCP / SYNCS) -----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// arrive and return the barrier state from BEFORE this arrival; mbar_pending_count(state) == 1 means this was the
// last pending arrival, i.e. the phase is now complete
__device__ __forceinline__ uint64_t mbar_arrive_state(uint64_t* bar) {
  uint64_t st;
  asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(smem_u32(bar)) : "memory");
  return st;
}
__device__ __forceinline__ uint32_t mbar_pending_count(uint64_t state) {
  uint32_t c;
  asm volatile("mbarrier.pending_count.b64 %0, %1;" : "=r"(c) : "l"(state));
  return c;
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe of a phase (try_wait may suspend the thread for a hardware-defined time)
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, both addresses 16-B aligned)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global, tracked by the per-thread bulk async-group
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---- thread-block clusters: CTA rank, peer shared-memory addresses, cluster barrier, remote store + signal ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA's window) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 8-byte store into a peer CTA's shared memory that completes 8 bytes of transaction on the peer's mbarrier: data and
// signal travel as one message (SASS: ST.ASYNC / SYNCS), the receiver waits on its own barrier
__device__ __forceinline__ void st_async_f64(uint32_t remote_addr, double v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr),
               "l"(__double_as_longlong(v)), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait_cluster(bar, parity)) {
  }
}

// streaming 16-byte load that does not pollute L1 (X is read once per pass)
__device__ __forceinline__ double2 ld_stream(const double2* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// 16-byte load of a small vector that every CTA re-reads (ts, u0): keep it in L1 against the streaming traffic
__device__ __forceinline__ double2 ld_keep(const double2* p) {
  double2 r;
  asm("ld.global.L1::evict_last.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

// streaming 16-byte store (written once, not re-read by this kernel)
__device__ __forceinline__ void st_stream(double2* p, const double2 v) {
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// block index of local feature j given B+1 ascending offsets (B is small)
__device__ __forceinline__ int block_of(const int* __restrict__ off, int B, int j) {
  int b = 0;
  while (b + 1 < B && j >= off[b + 1]) ++b;
  return b;
}

}  // namespace mbpls

// error codes of the C ABI (include/mbpls_b200.h)
#define MBPLS_OK 0
#define MBPLS_ERR_ARG (-1)
#define MBPLS_ERR_SIZE (-2)
#define MBPLS_CUDA_ERR(e) (1000 + static_cast<int>(e))
#define MBPLS_RETURN_LAST()                                  \
  do {                                                       \
    cudaError_t e__ = cudaGetLastError();                    \
    return e__ == cudaSuccess ? MBPLS_OK : MBPLS_CUDA_ERR(e__); \
  } while (0)
