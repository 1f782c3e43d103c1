// Building blocks of the SIMPLS / UNIPALS / KERNEL fit bodies (mbpls/mbpls.py:384-807, :995-1048) on the
// feature-major layout: multi-right-hand-side X'M, small vector algebra with fixed-order reductions, and
// the FP64 tensor-core (DMMA, mma.sync.m8n8k4.f64) cross-product kernel for X'X / XX'.
#include "launch.cuh"
#include "../../include/mbpls_b200.h"

using namespace mbpls;

// ------------------------------------------------------------------------------------------
// C[c][j] = sum_i Xt[j][i] * M[c][i]     (X'Y :998,:587; X'U :731; X'Ts :734; per-component X'Y of UNIPALS :396)
// One warp owns F=4 features and NC<=8 right-hand sides at a time: per step 4 X loads + NC M loads feed
// 8*NC FMAs, so the M traffic through L1 is at most 2x the HBM stream.  NaN entries of X count as zero.
// ------------------------------------------------------------------------------------------
template <int NC>
__global__ void __launch_bounds__(256)
xt_multi_kernel(const double* __restrict__ Xt, long ld, int n, int p, const double* __restrict__ M, long ldm, int c0, int C,
                double* __restrict__ out, long ldo, int chunk2, long part_stride) {
  // blockIdx.y = sample chunk (chunk2 16-byte units each; one chunk = the whole feature when gridDim.y == 1): with few, long
  // features (1 M samples x 2,000 features: 500 warp tasks) one warp per 4 whole features leaves the GPU at 0.5 TB/s, so the
  // sample axis is cut and the partial products of chunk s land in out + s * part_stride, to be added in chunk order
  constexpr int F = 4;
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nc = min(NC, C - c0);
  const int n2 = min(n >> 1, static_cast<int>(blockIdx.y + 1) * chunk2);
  const int i_begin = static_cast<int>(blockIdx.y) * chunk2;
  const bool tail_chunk = blockIdx.y == gridDim.y - 1;
  out += static_cast<size_t>(blockIdx.y) * part_stride;
  for (int j0 = gw * F; j0 < p; j0 += nwarps * F) {
    double acc[F][NC];
#pragma unroll
    for (int f = 0; f < F; ++f)
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[f][c] = 0.0;
    const double2* xr[F];
#pragma unroll
    for (int f = 0; f < F; ++f) xr[f] = reinterpret_cast<const double2*>(Xt + static_cast<size_t>(min(j0 + f, p - 1)) * ld);
    for (int i = i_begin + lane; i < n2; i += 32) {
      double2 x[F];
#pragma unroll
      for (int f = 0; f < F; ++f) {
        x[f] = ld_stream(xr[f] + i);
        if (isnan(x[f].x)) x[f].x = 0.0;
        if (isnan(x[f].y)) x[f].y = 0.0;
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (c < nc) {
          const double2 m = *reinterpret_cast<const double2*>(M + static_cast<size_t>(c0 + c) * ldm + 2 * i);
#pragma unroll
          for (int f = 0; f < F; ++f) {
            acc[f][c] = fma(x[f].x, m.x, acc[f][c]);
            acc[f][c] = fma(x[f].y, m.y, acc[f][c]);
          }
        }
      }
    }
    if ((n & 1) && lane == 0 && tail_chunk) {
#pragma unroll
      for (int f = 0; f < F; ++f) {
        double xv = Xt[static_cast<size_t>(min(j0 + f, p - 1)) * ld + n - 1];
        if (isnan(xv)) xv = 0.0;
#pragma unroll
        for (int c = 0; c < NC; ++c)
          if (c < nc) acc[f][c] = fma(xv, M[static_cast<size_t>(c0 + c) * ldm + n - 1], acc[f][c]);
      }
    }
#pragma unroll
    for (int f = 0; f < F; ++f)
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const double s = warp_sum(acc[f][c]);
        if (lane == 0 && c < nc && j0 + f < p) out[static_cast<size_t>(c0 + c) * ldo + j0 + f] = s;
      }
  }
}

// out[j] = base[j] - sum_k V[k][j] * coef[k]      (SIMPLS Gram-Schmidt steps, :1016-1017)
__global__ void __launch_bounds__(256)
lincomb_sub_kernel(double* __restrict__ out, const double* __restrict__ base, const double* __restrict__ V, long ldv, int K,
                   const double* __restrict__ coef, int len) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < len; j += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < K; ++k) s = fma(V[static_cast<size_t>(k) * ldv + j], coef[k], s);
    out[j] = base[j] - s;
  }
}

// t <- t - mean(t) (optional); nrm = ||t||; t <- t / nrm (optional).  Single CTA, fixed-order sums.  (:1007-1009)
__global__ void __launch_bounds__(1024)
center_normalize_kernel(double* __restrict__ t, int n, int center, int normalize, double* __restrict__ nrm_out) {
  __shared__ double scratch[32];
  double s = 0.0;
  if (center) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += t[i];
    s = block_sum1(s, scratch);
  }
  const double mean = center ? s / n : 0.0;
  double q = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = t[i] - mean;
    t[i] = v;
    q = fma(v, v, q);
  }
  q = block_sum1(q, scratch);
  const double nrm = sqrt(q);
  if (normalize)
    for (int i = threadIdx.x; i < n; i += blockDim.x) t[i] = t[i] / nrm;
  if (threadIdx.x == 0 && nrm_out) *nrm_out = nrm;
}

// out[b] = sum_{j in block b} w[j]^2   (a_b = ||w_b||^2, :408,:512,:609,:750); one CTA per block
__global__ void __launch_bounds__(256)
block_sumsq_kernel(const double* __restrict__ w, const int* __restrict__ off, double* __restrict__ out) {
  __shared__ double scratch[32];
  const int a = off[blockIdx.x], b = off[blockIdx.x + 1];
  double s = 0.0;
  for (int j = a + threadIdx.x; j < b; j += blockDim.x) s = fma(w[j], w[j], s);
  s = block_sum1(s, scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// out[j] = w[j] / sqrt(a[block(j)])   (w_b = partialloading / ||partialloading||, :407,:511,:607,:748)
__global__ void __launch_bounds__(256)
scale_by_block_kernel(const double* __restrict__ w, const int* __restrict__ off, int B, const double* __restrict__ a,
                      double* __restrict__ out, int p) {
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < p; j += gridDim.x * blockDim.x)
    out[j] = w[j] / sqrt(a[block_of(off, B, j)]);
}

// ------------------------------------------------------------------------------------------
// FP64 tensor-core cross product  C = op(A) . op(B)'  over the long dimension:
//   KMAJOR = true : C[i][j] = sum_k A[i*lda + k] * B[j*ldb + k]   (X'X from the feature-major matrix, :586)
//   KMAJOR = false: C[i][j] = sum_k A[k*lda + i] * B[k*ldb + j]   (XX' from the feature-major matrix, :704)
// 128x128 CTA tile, 16-deep k slabs in a 3-stage cp.async (LDGSTS) ring, 8 warps (2x4), each warp 64x32 =
// 8x4 mma.sync.aligned.m8n8k4.f64 tiles (64 accumulator doubles per lane).  Split-K over gridDim.z with
// per-split partial tiles reduced in fixed order by mbpls_reduce_chunks_f64 (deterministic).
// ------------------------------------------------------------------------------------------
#define XP_BM 128
#define XP_BN 128
#ifndef XP_BK
#define XP_BK 16  // 32 (221 KB of tiles) measured: X'X on 1 M x 2,000 156 -> 150 ms, XX' on 5,000 x 50,000 49 -> 54 ms; fragment
#endif            // prefetch across k-steps: no change (the compiler already hoists the loads)
#define XP_LD (XP_BK + 4)  // padded k-stride of a smem row (4 mod 16 -> conflict-free 64-bit fragment loads)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

#ifndef XP_STAGES
#define XP_STAGES 3
#endif

// cp.async (LDGSTS): asynchronous global -> shared copies of 8 / 16 bytes; src_bytes = 0 zero-fills the slot
__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool KMAJOR>
__global__ void __launch_bounds__(256)
crossprod_kernel(const double* __restrict__ A, long lda, const double* __restrict__ B, long ldb, int M, int N, long Kdim,
                 long k_per_split, double* __restrict__ Cpart, long ldc, int symmetric) {
  if (symmetric && blockIdx.x < blockIdx.y) return;  // SYRK: only tiles on / above the diagonal (mirrored afterwards)
  extern __shared__ __align__(16) double xp_smem[];
  constexpr int TILE = XP_BM * XP_LD;  // doubles per operand tile
  double* As = xp_smem;                      // [XP_STAGES][TILE]
  double* Bs = xp_smem + XP_STAGES * TILE;   // [XP_STAGES][TILE]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps
  const int m0 = blockIdx.y * XP_BM, n0 = blockIdx.x * XP_BN;
  const long kb = static_cast<long>(blockIdx.z) * k_per_split;
  const long ke = min(Kdim, kb + k_per_split);
  const int ntiles = kb < ke ? static_cast<int>((ke - kb + XP_BK - 1) / XP_BK) : 0;
  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  // asynchronous tile load: 128 rows x 16 k of A and of B -> smem [row][k] (k-stride XP_LD)
  auto issue_tile = [&](int slot, long k0) {
    double* as = As + slot * TILE;
    double* bs = Bs + slot * TILE;
    if (KMAJOR) {
#pragma unroll
      for (int e = 0; e < XP_BK / 4; ++e) {  // 16-byte copies along k (lda, ldb even; k0 even)
        const int idx = tid + e * 256;
        const int row = idx / (XP_BK / 2), kk = (idx % (XP_BK / 2)) * 2;
        const long k = k0 + kk;
        const int kbytes = k + 1 < ke ? 16 : (k < ke ? 8 : 0);
        const bool oka = (m0 + row < M) && kbytes > 0, okb = (n0 + row < N) && kbytes > 0;
        cp_async16(as + row * XP_LD + kk, oka ? A + static_cast<size_t>(m0 + row) * lda + k : A, oka ? kbytes : 0);
        cp_async16(bs + row * XP_LD + kk, okb ? B + static_cast<size_t>(n0 + row) * ldb + k : B, okb ? kbytes : 0);
      }
    } else {
#pragma unroll
      for (int e = 0; e < XP_BK / 2; ++e) {  // 8-byte copies, coalesced along the row (m / n) index
        const int idx = tid + e * 256;
        const int kk = idx >> 7, row = idx & 127;
        const long k = k0 + kk;
        const bool oka = (k < ke) && (m0 + row < M), okb = (k < ke) && (n0 + row < N);
        cp_async8(as + row * XP_LD + kk, oka ? A + static_cast<size_t>(k) * lda + m0 + row : A, oka ? 8 : 0);
        cp_async8(bs + row * XP_LD + kk, okb ? B + static_cast<size_t>(k) * ldb + n0 + row : B, okb ? 8 : 0);
      }
    }
  };

#pragma unroll
  for (int s = 0; s < XP_STAGES - 1; ++s) {
    if (s < ntiles) issue_tile(s, kb + static_cast<long>(s) * XP_BK);
    cp_async_commit();
  }
  for (int t = 0; t < ntiles; ++t) {
    cp_async_wait<XP_STAGES - 2>();  // tile t has landed (this thread's copies) ...
    __syncthreads();                 // ... and everybody's; also: everyone is done reading the slot refilled below
    const int tn = t + XP_STAGES - 1;
    if (tn < ntiles) issue_tile(tn % XP_STAGES, kb + static_cast<long>(tn) * XP_BK);
    cp_async_commit();
    const double* as = As + (t % XP_STAGES) * TILE + (wm * 64 + (lane >> 2)) * XP_LD + (lane & 3);
    const double* bs = Bs + (t % XP_STAGES) * TILE + (wn * 32 + (lane >> 2)) * XP_LD + (lane & 3);
#pragma unroll
    for (int ks = 0; ks < XP_BK; ks += 4) {
      double a[8], b[4];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = as[i * 8 * XP_LD + ks];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = bs[j * 8 * XP_LD + ks];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
  }
  cp_async_wait<0>();
  double* Cp = Cpart + static_cast<size_t>(blockIdx.z) * M * ldc;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = m0 + wm * 64 + i * 8 + (lane >> 2);
      const int c = n0 + wn * 32 + j * 8 + 2 * (lane & 3);
      if (r < M) {
        if (c < N) Cp[static_cast<size_t>(r) * ldc + c] = acc[i][j][0];
        if (c + 1 < N) Cp[static_cast<size_t>(r) * ldc + c + 1] = acc[i][j][1];
      }
    }
}

// C[j][i] = C[i][j] for the 128x128 tiles strictly below the diagonal (completes a symmetric crossprod)
__global__ void __launch_bounds__(256) symmetrize_kernel(double* __restrict__ C, long ldc, int M) {
  const int bi = blockIdx.y, bj = blockIdx.x;  // destination tile (row block bi, col block bj), bj < bi
  if (bj >= bi) return;
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int si = 0; si < XP_BM; si += 32)
    for (int sj = 0; sj < XP_BN; sj += 32) {
      // source element (r = bj*128+sj+.., c = bi*128+si+..) -> destination (c, r)
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = bj * XP_BM + sj + ty + 8 * k, c = bi * XP_BN + si + tx;
        if (r < M && c < M) tile[ty + 8 * k][tx] = C[static_cast<size_t>(r) * ldc + c];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = bi * XP_BN + si + ty + 8 * k, r = bj * XP_BM + sj + tx;
        if (r < M && c < M) C[static_cast<size_t>(c) * ldc + r] = tile[tx][ty + 8 * k];
      }
    }
}

// y[i] = sum_j A[i*lda + j] * x[j]  for a dense (symmetric) m x m matrix: one warp per row  (VAR w, AS_X ts)
__global__ void __launch_bounds__(256) dense_gemv_kernel(const double* __restrict__ A, long lda, int m, int ncols,
                                                         const double* __restrict__ x, double* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = gw; i < m; i += nwarps) {
    const double* a = A + static_cast<size_t>(i) * lda;
    double s = 0.0;
    for (int j = lane; j < ncols; j += 32) s = fma(a[j], x[j], s);
    s = warp_sum(s);
    if (lane == 0) y[i] = s;
  }
}

// A[i][j] += alpha * x[i]*y[j] + beta * y[i]*x[j] + gamma * x[i]*x[j]   (rank-1 / rank-2 forms of the sandwich
// deflations D'SD, D'VAR D (:630-633) and D AS D (:722-724))
__global__ void __launch_bounds__(256)
dense_rank2_kernel(double* __restrict__ A, long lda, int row0, int m, int ncols, const double* __restrict__ x,
                   const double* __restrict__ y, const double* __restrict__ scal, double alpha, double beta, double gamma,
                   int alpha_from, int gamma_from) {
  const int i = row0 + blockIdx.y;
  if (i >= m) return;
  // coefficients may be multiplied by a device scalar (e.g. -den or ts'K ts)
  const double al = alpha_from >= 0 ? alpha * scal[alpha_from] : alpha;
  const double be = alpha_from >= 0 ? beta * scal[alpha_from] : beta;
  const double ga = gamma_from >= 0 ? gamma * scal[gamma_from] : gamma;
  const double xi = x[i], yi = y[i];
  double* a = A + static_cast<size_t>(i) * lda;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < ncols; j += gridDim.x * blockDim.x)
    a[j] += al * xi * y[j] + be * yi * x[j] + ga * xi * x[j];
}

extern "C" {

int mbpls_xt_multi_f64(const double* Xt, long ld, int n, int p, const double* M, long ldm, int C, double* out, long ldo,
                       void* stream) {
  if (!Xt || !M || !out || C < 1 || (ld % 2) != 0 || (ldm % 2) != 0) return MBPLS_ERR_ARG;
  if (p == 0) return MBPLS_OK;
  int grid = (p + 31) / 32;  // 8 warps x 4 features
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int chunk2 = (n >> 1) + 1;
  for (int c0 = 0; c0 < C; c0 += 8) {
    const int nc = C - c0;
    if (nc > 4) xt_multi_kernel<8><<<grid, 256, 0, st>>>(Xt, ld, n, p, M, ldm, c0, C, out, ldo, chunk2, 0);
    else if (nc > 2) xt_multi_kernel<4><<<grid, 256, 0, st>>>(Xt, ld, n, p, M, ldm, c0, C, out, ldo, chunk2, 0);
    else if (nc > 1) xt_multi_kernel<2><<<grid, 256, 0, st>>>(Xt, ld, n, p, M, ldm, c0, C, out, ldo, chunk2, 0);
    else xt_multi_kernel<1><<<grid, 256, 0, st>>>(Xt, ld, n, p, M, ldm, c0, C, out, ldo, chunk2, 0);
  }
  MBPLS_RETURN_LAST();
}

/* sample chunks worth cutting the features of an X'M product into (1: use mbpls_xt_multi_f64) */
int mbpls_xt_multi_chunks(int n, int p) {
  if (p <= 0 || n < (1 << 16)) return 1;
  const long tasks = (static_cast<long>(p) + 3) / 4;          // warp tasks without a cut
  const long want = static_cast<long>(num_sms()) * 8 * 6;     // ~6 tasks per resident warp
  long s = (want + tasks - 1) / tasks;
  const long smax = n / (1 << 14);                            // chunks of at least 16k samples
  if (s > smax) s = smax;
  if (s > 256) s = 256;
  return s < 1 ? 1 : static_cast<int>(s);
}

/* the same product with every feature cut into `chunks` sample ranges: part[s][c][j] (part_stride = C * ldo doubles per
 * chunk) holds the partial products of chunk s; add them in chunk order (mbpls_reduce_chunks_f64(part, chunks, C * ldo, out)) */
int mbpls_xt_multi_split_f64(const double* Xt, long ld, int n, int p, const double* M, long ldm, int C, double* part, long ldo,
                             int chunks, void* stream) {
  if (!Xt || !M || !part || C < 1 || chunks < 1 || chunks > 65535 || (ld % 2) != 0 || (ldm % 2) != 0) return MBPLS_ERR_ARG;
  if (p == 0) return MBPLS_OK;
  int gx = (p + 31) / 32;  // 8 warps x 4 features
  if (gx > num_sms() * 8) gx = num_sms() * 8;
  const dim3 grid(gx, chunks);
  const int n2 = n >> 1;
  int chunk2 = (n2 + chunks - 1) / chunks;
  chunk2 = (chunk2 + 31) / 32 * 32;
  if (chunk2 < 32) chunk2 = 32;
  const long stride = static_cast<long>(C) * ldo;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int c0 = 0; c0 < C; c0 += 8) {
    const int nc = C - c0;
    if (nc > 4) xt_multi_kernel<8><<<grid, 256, 0, st>>>(Xt, ld, n, p, M, ldm, c0, C, part, ldo, chunk2, stride);
    else if (nc > 2) xt_multi_kernel<4><<<grid, 256, 0, st>>>(Xt, ld, n, p, M, ldm, c0, C, part, ldo, chunk2, stride);
    else if (nc > 1) xt_multi_kernel<2><<<grid, 256, 0, st>>>(Xt, ld, n, p, M, ldm, c0, C, part, ldo, chunk2, stride);
    else xt_multi_kernel<1><<<grid, 256, 0, st>>>(Xt, ld, n, p, M, ldm, c0, C, part, ldo, chunk2, stride);
  }
  MBPLS_RETURN_LAST();
}

int mbpls_lincomb_sub_f64(double* out, const double* base, const double* V, long ldv, int K, const double* coef, int len,
                          void* stream) {
  if (!out || !base || (K > 0 && (!V || !coef))) return MBPLS_ERR_ARG;
  if (len <= 0) return MBPLS_OK;
  int grid = (len + 255) / 256;
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  lincomb_sub_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, base, V, ldv, K, coef, len);
  MBPLS_RETURN_LAST();
}

int mbpls_center_normalize_f64(double* t, int n, int center, int normalize, double* nrm_out, void* stream) {
  if (!t || n < 1) return MBPLS_ERR_ARG;
  center_normalize_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(t, n, center, normalize, nrm_out);
  MBPLS_RETURN_LAST();
}

int mbpls_block_sumsq_f64(const double* w, const int* off, int B, double* out, void* stream) {
  if (!w || !off || !out || B < 1) return MBPLS_ERR_ARG;
  block_sumsq_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, off, out);
  MBPLS_RETURN_LAST();
}

int mbpls_scale_by_block_f64(const double* w, const int* off, int B, const double* a, double* out, int p, void* stream) {
  if (!w || !off || !a || !out || B < 1) return MBPLS_ERR_ARG;
  if (p <= 0) return MBPLS_OK;
  int grid = (p + 255) / 256;
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  scale_by_block_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, off, B, a, out, p);
  MBPLS_RETURN_LAST();
}

int mbpls_crossprod_splits(int M, int N, long Kdim) {
  const long tiles = static_cast<long>((M + XP_BM - 1) / XP_BM) * ((N + XP_BN - 1) / XP_BN);
  long want = (2L * num_sms() + tiles - 1) / tiles;  // ~2 waves of CTAs
  const long maxs = (Kdim + 4 * XP_BK - 1) / (4 * XP_BK);
  if (want > maxs) want = maxs;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  return static_cast<int>(want);
}

// Split count of a SYMMETRIC cross product (A == B, M == N): only the nb (nb + 1) / 2 tiles on / above the diagonal do work and
// every CTA owns a whole SM (123 KB of tiles), so the launch takes ceil(working CTAs / SMs) rounds of one CTA's k range.  With the
// rectangular rule above X'X of 1 M x 2,000 runs 272 working CTAs on 148 SMs (two rounds, 8 % of the SM-time idle: ncu shows the
// FP64 tensor pipe 84 % busy while active but 77 % of the elapsed time, 42 % on the least loaded SM).  Here the count minimises
// rounds x k-slabs per CTA (+ a few slabs of pipeline fill / epilogue per round) + the traffic of the partial matrices, under
// a 1 GiB cap on the partials; the choice depends on the shape and the SM count only, so it is the same on every rank and run.
int mbpls_crossprod_splits_syrk(int M, long Kdim) {
  if (M < 1 || Kdim < 1) return 1;
  const long nb = (M + XP_BM - 1) / XP_BM, working = nb * (nb + 1) / 2, nsm = num_sms() > 0 ? num_sms() : 148;
  long maxs = (Kdim + 4 * XP_BK - 1) / (4 * XP_BK);
  const long cap = (1L << 30) / (8L * M * ((M + 15) / 16 * 16));
  if (maxs > cap) maxs = cap;
  if (maxs > 64) maxs = 64;
  if (maxs < 1) maxs = 1;
  const double slab_us = 2.6;                                         // one 128 x 128 x 16 slab on one SM at ~0.2 TFLOP/s
  const double part_us = 2.5 * 8.0 * M * static_cast<double>(M) / 6.0e6;  // zero-fill + write + read of one partial at 6 TB/s
  long best = 1;
  double best_us = 0.0;
  for (long s = 1; s <= maxs; ++s) {
    const long rounds = (working * s + nsm - 1) / nsm;
    const long slabs = ((Kdim + s - 1) / s + XP_BK - 1) / XP_BK;
    const double us = rounds * (slabs + 4) * slab_us + s * part_us;
    if (s == 1 || us < best_us * (1.0 - 1e-3)) {  // a later count has to win by more than 0.1 %
      best = s;
      best_us = us;
    }
  }
  return static_cast<int>(best);
}

// kmajor = 1: C = A B' with A (M x Kdim, lda), B (N x Kdim, ldb); kmajor = 0: C = A' B with A (Kdim x M), B (Kdim x N).
// Cpart holds `splits` partial M x ldc matrices (splits = mbpls_crossprod_splits); reduce with mbpls_reduce_chunks_f64.
int mbpls_symmetrize_f64(double* C, long ldc, int M, void* stream) {
  if (!C || M < 1 || ldc < M) return MBPLS_ERR_ARG;
  const int nb = (M + XP_BM - 1) / XP_BM;
  if (nb < 2) return MBPLS_OK;
  dim3 grid(nb, nb);
  symmetrize_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(C, ldc, M);
  MBPLS_RETURN_LAST();
}

int mbpls_crossprod_f64(const double* A, long lda, const double* B, long ldb, int M, int N, long Kdim, int kmajor, int splits,
                        double* Cpart, long ldc, int symmetric, void* stream) {
  if (symmetric && (A != B || M != N)) return MBPLS_ERR_ARG;
  if (!A || !B || !Cpart || M < 1 || N < 1 || Kdim < 0 || splits < 1 || ldc < N) return MBPLS_ERR_ARG;
  if (kmajor && ((lda % 2) != 0 || (ldb % 2) != 0)) return MBPLS_ERR_ARG;
  long kps = (Kdim + splits - 1) / splits;
  kps = ((kps + XP_BK - 1) / XP_BK) * XP_BK;
  if (kps < XP_BK) kps = XP_BK;
  dim3 grid((N + XP_BN - 1) / XP_BN, (M + XP_BM - 1) / XP_BM, splits);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int smem = XP_STAGES * (XP_BM + XP_BN) * XP_LD * static_cast<int>(sizeof(double));
  if (kmajor) {
    cudaFuncSetAttribute(crossprod_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    crossprod_kernel<true><<<grid, 256, smem, st>>>(A, lda, B, ldb, M, N, Kdim, kps, Cpart, ldc, symmetric);
  } else {
    cudaFuncSetAttribute(crossprod_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    crossprod_kernel<false><<<grid, 256, smem, st>>>(A, lda, B, ldb, M, N, Kdim, kps, Cpart, ldc, symmetric);
  }
  MBPLS_RETURN_LAST();
}

int mbpls_dense_gemv_f64(const double* A, long lda, int m, int ncols, const double* x, double* y, void* stream) {
  if (!A || !x || !y || m < 1) return MBPLS_ERR_ARG;
  int grid = (m + 7) / 8;
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  dense_gemv_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(A, lda, m, ncols, x, y);
  MBPLS_RETURN_LAST();
}

int mbpls_dense_rank2_f64(double* A, long lda, int m, int ncols, const double* x, const double* y, const double* scal,
                          double alpha, double beta, double gamma, int alpha_from, int gamma_from, void* stream) {
  if (!A || !x || !y || m < 1) return MBPLS_ERR_ARG;
  int gx = (ncols + 255) / 256;
  if (gx > 32) gx = 32;
  for (int i0 = 0; i0 < m; i0 += 65535) {
    const int mi = (m - i0) < 65535 ? (m - i0) : 65535;
    dim3 grid(gx, mi);
    dense_rank2_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(A, lda, i0, m, ncols, x, y, scal, alpha, beta,
                                                                           gamma, alpha_from, gamma_from);
  }
  MBPLS_RETURN_LAST();
}

}  // extern "C"
