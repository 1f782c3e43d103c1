// Small dense linear algebra of the per-component steps, on the device (single CTA, shared memory):
//   top eigenvector of a q x q symmetric PSD matrix        (np.linalg.svd(S)[0][:, 0:1] with S = C C', mbpls.py:398,590,1001)
//   top left singular vector of A B' from G = B'B, H = A'A   (np.linalg.svd(XX'YY')[0][:, 0:1], :491, :713)
//   Moore-Penrose pseudo-inverse of a K x K matrix           (np.linalg.pinv(P'W), :476, :569, :642, :737, :988; (Ts'Ts)^+ :734)
// All three are built on one primitive: one-sided (Hestenes) Jacobi SVD  W V = U Sigma  with the column pairs of
// each round-robin round rotated by one warp each (up to 32 disjoint pairs -> m <= 64).  One-sided Jacobi works on
// the matrix itself (no M'M squaring) and delivers singular vectors to high relative accuracy.
#include "launch.cuh"
#include "../../include/mbpls_b200.h"

using namespace mbpls;

#define SL_MAX 64
#define SL_LD 65  // column stride in shared memory (column-major, padded)

// W, V: column-major m x m in shared memory (element (r, c) at [c * SL_LD + r]).  On exit the columns of W are
// U_j * sigma_j and V holds the right singular vectors.  Must be called by all 1024 threads of the CTA.
__device__ void onesided_jacobi(double* W, double* V, int m, int* s_flag) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int me = m + (m & 1), npairs = me >> 1;
  for (int e = tid; e < m * m; e += blockDim.x) {
    const int c = e / m, r = e % m;
    V[c * SL_LD + r] = (r == c) ? 1.0 : 0.0;
  }
  __syncthreads();
  for (int sweep = 0; sweep < 60; ++sweep) {
    if (tid == 0) *s_flag = 0;
    __syncthreads();
    for (int round = 0; round < me - 1; ++round) {
      if (warp < npairs) {
        int p, q;
        if (warp == 0) {
          p = me - 1;
          q = round;
        } else {
          p = (round + warp) % (me - 1);
          q = (round - warp + (me - 1)) % (me - 1);
        }
        if (p > q) {
          const int t = p;
          p = q;
          q = t;
        }
        if (q < m) {  // the padding column of an odd m takes no part
          double* wp = W + p * SL_LD;
          double* wq = W + q * SL_LD;
          double al = 0.0, be = 0.0, ga = 0.0;
          for (int i = lane; i < m; i += 32) {
            const double a = wp[i], b = wq[i];
            al = fma(a, a, al);
            be = fma(b, b, be);
            ga = fma(a, b, ga);
          }
          al = warp_sum(al);
          be = warp_sum(be);
          ga = warp_sum(ga);
          if (fabs(ga) > 1e-300 && fabs(ga) > 1.1e-16 * sqrt(al * be)) {
            const double zeta = (be - al) / (2.0 * ga);
            const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
            const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
            double* vp = V + p * SL_LD;
            double* vq = V + q * SL_LD;
            for (int i = lane; i < m; i += 32) {
              const double a = wp[i], b = wq[i];
              wp[i] = c * a - s * b;
              wq[i] = s * a + c * b;
              const double x = vp[i], y = vq[i];
              vp[i] = c * x - s * y;
              vq[i] = s * x + c * y;
            }
            if (lane == 0) *s_flag = 1;
          }
        }
      }
      __syncthreads();
    }
    const int rotated = *s_flag;
    __syncthreads();
    if (!rotated) break;
  }
}

// sigma[j] = ||W[:, j]||; returns (to every thread) the index of the largest one
__device__ int column_norms(const double* W, int m, double* sigma) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int j = warp; j < m; j += blockDim.x >> 5) {
    double s = 0.0;
    for (int i = lane; i < m; i += 32) s = fma(W[j * SL_LD + i], W[j * SL_LD + i], s);
    s = warp_sum(s);
    if (lane == 0) sigma[j] = sqrt(s);
  }
  __syncthreads();
  int best = 0;
  for (int j = 1; j < m; ++j)
    if (sigma[j] > sigma[best]) best = j;
  return best;
}

__device__ void load_colmajor(double* dst, const double* __restrict__ src, long ld, int m, bool symmetrize) {
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int r = e / m, c = e % m;
    double v = src[static_cast<size_t>(r) * ld + c];
    if (symmetrize) v = 0.5 * (v + src[static_cast<size_t>(c) * ld + r]);
    dst[c * SL_LD + r] = v;
  }
  __syncthreads();
}

// out[0..m) = unit eigenvector of the largest eigenvalue of the symmetric PSD matrix G (m x m, row-major, ld)
__global__ void __launch_bounds__(1024) small_top_eigvec_kernel(const double* __restrict__ G, long ld, int m, double* __restrict__ out) {
  extern __shared__ __align__(16) double sl_smem[];
  double *W = sl_smem, *V = sl_smem + SL_MAX * SL_LD, *sigma = sl_smem + 2 * SL_MAX * SL_LD;
  __shared__ int flag;
  load_colmajor(W, G, ld, m, true);
  onesided_jacobi(W, V, m, &flag);
  const int best = column_norms(W, m, sigma);
  for (int i = threadIdx.x; i < m; i += blockDim.x) out[i] = V[best * SL_LD + i];
}

// out (m x m, row-major, ldo) = pinv(M), singular values <= rcond * sigma_max treated as zero (numpy semantics)
__global__ void __launch_bounds__(1024)
small_pinv_kernel(const double* __restrict__ M, long ld, int m, double rcond, double* __restrict__ out, long ldo) {
  extern __shared__ __align__(16) double sl_smem[];
  double *W = sl_smem, *V = sl_smem + SL_MAX * SL_LD, *sigma = sl_smem + 2 * SL_MAX * SL_LD;
  __shared__ int flag;
  load_colmajor(W, M, ld, m, false);
  onesided_jacobi(W, V, m, &flag);
  const int best = column_norms(W, m, sigma);
  const double cut = rcond * sigma[best];
  // pinv = V Sigma^+ U' with U_k = W[:, k] / sigma_k  ->  out[i][j] = sum_k V[i][k] * W[j][k] / sigma_k^2
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int i = e / m, j = e % m;
    double s = 0.0;
    for (int k = 0; k < m; ++k) {
      const double sk = sigma[k];
      if (sk > cut) s += V[k * SL_LD + i] * (W[k * SL_LD + j] / sk) / sk;
    }
    out[static_cast<size_t>(i) * ldo + j] = s;
  }
}

// c (length m) such that A c is the top left singular vector of A B', from G = B'B and H = A'A (both m x m):
// G = L L' with L = Q diag(sqrt(lambda)); z = top eigenvector of L' H L; c = L z.
__global__ void __launch_bounds__(1024)
small_top_sv_product_kernel(const double* __restrict__ G, long ldg, const double* __restrict__ H, long ldh, int m,
                            double* __restrict__ out) {
  extern __shared__ __align__(16) double sl_smem[];
  double *W = sl_smem, *V = sl_smem + SL_MAX * SL_LD, *sigma = sl_smem + 2 * SL_MAX * SL_LD,
         *L = sl_smem + 2 * SL_MAX * SL_LD + SL_MAX;
  __shared__ int flag;
  load_colmajor(W, G, ldg, m, true);
  onesided_jacobi(W, V, m, &flag);   // G V = V diag(lambda): columns of W have norm lambda_k
  column_norms(W, m, sigma);
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int k = e / m, r = e % m;
    L[k * SL_LD + r] = V[k * SL_LD + r] * sqrt(sigma[k]);  // L[:, k] = q_k sqrt(lambda_k)
  }
  __syncthreads();
  // T = H L (columns), then M2 = L' T, both m x m
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int k = e / m, r = e % m;
    double s = 0.0;
    for (int j = 0; j < m; ++j) s = fma(0.5 * (H[static_cast<size_t>(r) * ldh + j] + H[static_cast<size_t>(j) * ldh + r]), L[k * SL_LD + j], s);
    V[k * SL_LD + r] = s;  // V is free again: holds T
  }
  __syncthreads();
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int c = e / m, r = e % m;
    double s = 0.0;
    for (int j = 0; j < m; ++j) s = fma(L[r * SL_LD + j], V[c * SL_LD + j], s);
    W[c * SL_LD + r] = s;  // M2[r][c]
  }
  __syncthreads();
  // symmetrise M2 in place (rounding), then top eigenvector
  for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
    const int c = e / m, r = e % m;
    if (r < c) {
      const double v = 0.5 * (W[c * SL_LD + r] + W[r * SL_LD + c]);
      W[c * SL_LD + r] = v;
      W[r * SL_LD + c] = v;
    }
  }
  __syncthreads();
  onesided_jacobi(W, V, m, &flag);
  const int best = column_norms(W, m, sigma);
  for (int i = threadIdx.x; i < m; i += blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < m; ++k) s = fma(L[k * SL_LD + i], V[best * SL_LD + k], s);
    out[i] = s;
  }
}

extern "C" {

int mbpls_small_top_eigvec_f64(const double* G, long ld, int m, double* out, void* stream) {
  if (!G || !out || m < 1) return MBPLS_ERR_ARG;
  if (m > SL_MAX) return MBPLS_ERR_SIZE;
  const int smem = (2 * SL_MAX * SL_LD + SL_MAX) * static_cast<int>(sizeof(double));
  cudaFuncSetAttribute(small_top_eigvec_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  small_top_eigvec_kernel<<<1, 1024, smem, static_cast<cudaStream_t>(stream)>>>(G, ld, m, out);
  MBPLS_RETURN_LAST();
}

int mbpls_small_pinv_f64(const double* M, long ld, int m, double rcond, double* out, long ldo, void* stream) {
  if (!M || !out || m < 1) return MBPLS_ERR_ARG;
  if (m > SL_MAX) return MBPLS_ERR_SIZE;
  const int smem = (2 * SL_MAX * SL_LD + SL_MAX) * static_cast<int>(sizeof(double));
  cudaFuncSetAttribute(small_pinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  small_pinv_kernel<<<1, 1024, smem, static_cast<cudaStream_t>(stream)>>>(M, ld, m, rcond, out, ldo);
  MBPLS_RETURN_LAST();
}

int mbpls_small_top_sv_product_f64(const double* G, long ldg, const double* H, long ldh, int m, double* out, void* stream) {
  if (!G || !H || !out || m < 1) return MBPLS_ERR_ARG;
  if (m > SL_MAX) return MBPLS_ERR_SIZE;
  const int smem = (3 * SL_MAX * SL_LD + SL_MAX) * static_cast<int>(sizeof(double));
  cudaFuncSetAttribute(small_top_sv_product_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  small_top_sv_product_kernel<<<1, 1024, smem, static_cast<cudaStream_t>(stream)>>>(G, ldg, H, ldh, m, out);
  MBPLS_RETURN_LAST();
}

}  // extern "C"
