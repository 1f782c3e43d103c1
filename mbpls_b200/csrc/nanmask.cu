// NaN pattern as a bit matrix, and the masked denominators of the NaN-mode NIPALS (mbpls/mbpls.py:848-852, :867-872,
// :923-925) computed from it.
//
// The NaN pattern of X never changes during a fit (deflation keeps NaN as NaN, :969), so it is extracted once:
// bits[j][i >> 5] bit (i & 31) = 1 iff x_ij is NaN (1/64 of the size of X).  With it, every masked sum of the
// reference becomes "total minus what sits under the holes":
//   * per feature:  sum_{i observed in column j} v_i^2 = v'v - sum_{i: x_ij NaN} v_i^2      (v = u, ts, u0)
//   * per sample:   sum_{j in block, observed in row i} w_j^2                                 (score denominators)
// Both are tiny passes over the bit matrix instead of full passes over X, which is what lets the NaN-mode trip run
// through the same one-read kernel as dense data (csrc/fused.cu) with NaN entries read as zero.
#include "launch.cuh"
#include "../../include/mbpls_b200.h"

using namespace mbpls;

namespace {

// one warp per feature; lanes read consecutive samples, so a ballot is the mask word
__global__ void __launch_bounds__(256) nan_bitmask_kernel(const double* __restrict__ Xt, long ld, int n, int p,
                                                          unsigned* __restrict__ bits, long ldw) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int nwords = (n + 31) >> 5;
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < p; j += warps) {
    const double* x = Xt + static_cast<size_t>(j) * ld;
    unsigned* row = bits + static_cast<size_t>(j) * ldw;
    for (int w = 0; w < nwords; w += 4) {  // four independent loads in flight per lane
      double v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = (w + k) * 32 + lane;
        v[k] = i < n ? x[i] : 0.0;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const unsigned m = __ballot_sync(MBPLS_FULL_MASK, isnan(v[k]));
        if (lane == 0 && w + k < ldw) row[w + k] = m;
      }
    }
    for (int w = ((nwords + 3) & ~3) + lane; w < ldw; w += 32) row[w] = 0u;
  }
}

// m_j = v . v2 - sum_{i: x_ij NaN} v_i v2_i, the sum of v_i v2_i over the observed samples of feature j (fully observed
// features: v . v2).  mode 0: 1/m_j for features with NaN, 1 otherwise (the reference does not divide dense loadings,
// :920); mode 1: 1/m_j; mode 2: m_j.
// SMEM: the products v_i v2_i are staged once per CTA in shared memory (8 n bytes), so the gather under every set bit is a
// shared-memory read instead of a global one (ncu on the global form: 64 % of the stall samples on those gathers).
template <bool SMEM>
__global__ void __launch_bounds__(SMEM ? 1024 : 256, SMEM ? 2 : 1) masked_colden_kernel(const unsigned* __restrict__ bits, long ldw, int n, int p,
                                                            const int* __restrict__ col_nan, const double* __restrict__ v,
                                                            const double* __restrict__ v2, const double* __restrict__ vv_ptr, int mode,
                                                            double* __restrict__ out, const int* __restrict__ done) {
  if (done && *done) return;
  extern __shared__ double cd_prod[];
  if (SMEM) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) cd_prod[i] = v[i] * v2[i];
    __syncthreads();
  }
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int nwords = (n + 31) >> 5;
  const double vv = *vv_ptr;
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < p; j += warps) {
    if (col_nan[j] == 0) {
      if (lane == 0) out[j] = mode == 0 ? 1.0 : (mode == 1 ? 1.0 / vv : vv);
      continue;
    }
    const unsigned* row = bits + static_cast<size_t>(j) * ldw;
    // two mask words per lane at a time, each with its own accumulator: two independent gather + add chains per iteration, and
    // the iteration count of the warp (its longest lane) follows max(popc) over word pairs instead of the sum over single words
    double miss = 0.0, miss2 = 0.0;
    for (int w = lane; w < nwords; w += 64) {
      unsigned ma = row[w];
      unsigned mb = w + 32 < nwords ? row[w + 32] : 0u;
      const int ia = w * 32, ib = ia + 1024;
      while (ma | mb) {
        if (ma) {
          const int b = __ffs(ma) - 1;
          ma &= ma - 1;
          miss += SMEM ? cd_prod[ia + b] : v[ia + b] * v2[ia + b];
        }
        if (mb) {
          const int b = __ffs(mb) - 1;
          mb &= mb - 1;
          miss2 += SMEM ? cd_prod[ib + b] : v[ib + b] * v2[ib + b];
        }
      }
    }
    miss = warp_sum(miss + miss2);
    double obs = vv - miss;
    if (fabs(miss) > 0.5 * fabs(vv)) {
      // most of the mass sits under the holes: "total minus missing" would cancel (and could reach 0 or change sign where
      // the reference's direct sum over the observed samples, :849-852 / :923-925, is positive) -> sum the observed ones
      double o = 0.0;
      for (int w = lane; w < nwords; w += 32) {
        unsigned m = ~row[w];
        if ((w + 1) * 32 > n) m &= (n - w * 32 >= 32) ? 0xffffffffu : ((1u << (n - w * 32)) - 1u);
        while (m) {
          const int b = __ffs(m) - 1;
          m &= m - 1;
          o += SMEM ? cd_prod[w * 32 + b] : v[w * 32 + b] * v2[w * 32 + b];
        }
      }
      obs = warp_sum(o);
    }
    if (lane == 0) out[j] = mode == 2 ? obs : 1.0 / obs;
  }
}

__global__ void __launch_bounds__(1024) vec_dot_kernel(const double* __restrict__ a, const double* __restrict__ b, int n,
                                                       double* __restrict__ out) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc = fma(a[i], b[i], acc);
  acc = block_sum1(acc, scratch);
  if (threadIdx.x == 0) out[0] = acc;
}

// 1.0 where bit b of the mask word is clear (sample observed), 0.0 where it is set: only the high word depends on the bit (one
// 32-bit select), and q * {0.0, 1.0} + acc is exact, so the predicated addition `acc += bit ? 0 : q` (LOP3 + 2 FSEL + DADD) becomes
// LOP3 + SEL + DFMA -- the kernel is bound by its instruction count.
__device__ __forceinline__ double obs_factor(unsigned m, int b) {
  return __hiloint2double(((m >> b) & 1u) ? 0 : 0x3ff00000, 0);
}

// Tden[s][i] = sum over the features j of split s that are observed in sample i of w_j^2.
// A lane owns one 32-bit mask word = 32 consecutive samples and keeps their 32 sums in registers, so one 4-byte load feeds
// 32 predicated additions; a warp covers 1024 consecutive samples (its lanes read 128 contiguous bytes of a feature's bit
// row), the 8 warps of a CTA take every 8th feature of the split and their partials are added in a fixed order.
__global__ void __launch_bounds__(256) masked_rowden_kernel(const unsigned* __restrict__ bits, long ldw, int n,
                                                            const double* __restrict__ w, const int* __restrict__ split_f0,
                                                            const int* __restrict__ split_f1, double* __restrict__ Tden, long ldt,
                                                            const int* __restrict__ done) {
  if (done && *done) return;
  __shared__ double part[8][16][33];  // [warp][bit (half)][lane], padded against bank conflicts
  const int s = blockIdx.y;
  const int f0 = split_f0[s], f1 = split_f1[s];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wi = blockIdx.x * 32 + lane;
  const bool valid = wi < ldw;
  const unsigned* col = bits + (valid ? wi : 0);
  double acc[32];
#pragma unroll
  for (int b = 0; b < 32; ++b) acc[b] = 0.0;
  int j = f0 + warp;
  for (; j + 24 < f1; j += 32) {  // four features per step: all loads in flight before the additions
    const unsigned m0 = valid ? __ldg(col + static_cast<size_t>(j) * ldw) : 0u;
    const unsigned m1 = valid ? __ldg(col + static_cast<size_t>(j + 8) * ldw) : 0u;
    const unsigned m2 = valid ? __ldg(col + static_cast<size_t>(j + 16) * ldw) : 0u;
    const unsigned m3 = valid ? __ldg(col + static_cast<size_t>(j + 24) * ldw) : 0u;
    const double a0 = __ldg(w + j), a1 = __ldg(w + j + 8), a2 = __ldg(w + j + 16), a3 = __ldg(w + j + 24);
    const double q0 = a0 * a0, q1 = a1 * a1, q2 = a2 * a2, q3 = a3 * a3;
#pragma unroll
    for (int b = 0; b < 32; ++b) {
      acc[b] = fma(q0, obs_factor(m0, b), acc[b]);
      acc[b] = fma(q1, obs_factor(m1, b), acc[b]);
      acc[b] = fma(q2, obs_factor(m2, b), acc[b]);
      acc[b] = fma(q3, obs_factor(m3, b), acc[b]);
    }
  }
  for (; j + 8 < f1; j += 16) {  // two features per step
    const unsigned m0 = valid ? __ldg(col + static_cast<size_t>(j) * ldw) : 0u;
    const unsigned m1 = valid ? __ldg(col + static_cast<size_t>(j + 8) * ldw) : 0u;
    const double a0 = __ldg(w + j), a1 = __ldg(w + j + 8);
    const double q0 = a0 * a0, q1 = a1 * a1;
#pragma unroll
    for (int b = 0; b < 32; ++b) {
      acc[b] = fma(q0, obs_factor(m0, b), acc[b]);
      acc[b] = fma(q1, obs_factor(m1, b), acc[b]);
    }
  }
  for (; j < f1; j += 8) {
    const unsigned m0 = valid ? __ldg(col + static_cast<size_t>(j) * ldw) : 0u;
    const double a0 = __ldg(w + j);
    const double q0 = a0 * a0;
#pragma unroll
    for (int b = 0; b < 32; ++b) acc[b] = fma(q0, obs_factor(m0, b), acc[b]);
  }
  double* out = Tden + static_cast<size_t>(s) * ldt;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    __syncthreads();
#pragma unroll
    for (int b = 0; b < 16; ++b) part[warp][b][lane] = acc[half * 16 + b];
    __syncthreads();
    // 16 bits x 32 lanes = 512 sums per half, 256 threads: two each, warps 0..7 added in order
    for (int e = threadIdx.x; e < 512; e += 256) {
      const int b = e & 15, l = e >> 4;
      double t = 0.0;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += part[k][b][l];
      const long i = (static_cast<long>(blockIdx.x) * 32 + l) * 32 + half * 16 + b;
      if (i < n) out[i] = t;
    }
  }
}

}  // namespace

extern "C" {

/* words per feature row of the NaN bit matrix for n samples (multiple of 4: rows are 16-byte aligned) */
int mbpls_nan_bitmask_ldw(int n) { return (((n + 31) >> 5) + 3) & ~3; }

int mbpls_nan_bitmask_f64(const double* Xt, long ld, int n, int p, unsigned* bits, long ldw, void* stream) {
  if (!Xt || !bits || ld < n || ldw < ((n + 31) >> 5)) return MBPLS_ERR_ARG;
  if (p == 0 || n == 0) return MBPLS_OK;
  int grid = (p + 7) / 8;
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  nan_bitmask_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(Xt, ld, n, p, bits, ldw);
  MBPLS_RETURN_LAST();
}

int mbpls_masked_colden_f64(const unsigned* bits, long ldw, int n, int p, const int* col_nan, const double* v, const double* v2,
                            const double* vv, int mode, double* out, const int* done, void* stream) {
  if (!bits || !col_nan || !v || !vv || !out || mode < 0 || mode > 2) return MBPLS_ERR_ARG;
  if (p == 0) return MBPLS_OK;
  int grid = (p + 7) / 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t smem = static_cast<size_t>(n) * sizeof(double);
  if (smem <= 100 * 1024) {  // two 1024-thread CTAs per SM (all 64 warps: the gathers are latency-bound), each with its copy of the products
    grid = (p + 31) / 32;
    if (grid > num_sms() * 2) grid = num_sms() * 2;
    cudaFuncSetAttribute(masked_colden_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    masked_colden_kernel<true><<<grid, 1024, smem, st>>>(bits, ldw, n, p, col_nan, v, v2 ? v2 : v, vv, mode, out, done);
  } else {
    if (grid > num_sms() * 16) grid = num_sms() * 16;
    masked_colden_kernel<false><<<grid, 256, 0, st>>>(bits, ldw, n, p, col_nan, v, v2 ? v2 : v, vv, mode, out, done);
  }
  MBPLS_RETURN_LAST();
}

int mbpls_vec_dot_f64(const double* a, const double* b, int n, double* out, void* stream) {
  if (!a || !b || !out || n < 0) return MBPLS_ERR_ARG;
  vec_dot_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(a, b, n, out);
  MBPLS_RETURN_LAST();
}

int mbpls_masked_rowden_f64(const unsigned* bits, long ldw, int n, const double* w, const int* split_f0, const int* split_f1,
                            int nsplit, double* Tden, long ldt, const int* done, void* stream) {
  if (!bits || !w || !split_f0 || !split_f1 || !Tden) return MBPLS_ERR_ARG;
  if (nsplit == 0 || n == 0) return MBPLS_OK;
  if (nsplit > 65535) return MBPLS_ERR_SIZE;
  dim3 grid((((n + 31) >> 5) + 31) / 32, nsplit);  // 32 mask words (1024 samples) per CTA
  masked_rowden_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(bits, ldw, n, w, split_f0, split_f1, Tden, ldt, done);
  MBPLS_RETURN_LAST();
}

}  // extern "C"
