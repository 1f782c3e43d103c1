// Finalisation and new-data kernels:
//   R_ = W pinv(P'W), beta_ = R_ V_'                       (mbpls/mbpls.py:986-989, :642-643, :737-738)
//   Ts = X R_, y_hat = X beta_, T_b = X_b W_b               (:1110-1117, :1142-1155, :1379-1386)
// All small-matrix contractions over the feature axis are two-stage (per-chunk partial, fixed-order
// sum) so results are bitwise reproducible.
#include "launch.cuh"
#include "../../include/mbpls_b200.h"

using namespace mbpls;

#define GRAM_CHUNK 2048
#define GRAM_TILE 32

// Cpart[chunk][i*K2+j] = sum_{f in chunk} A[i][f] * Bm[j][f]
__global__ void __launch_bounds__(256)
gram_partial_kernel(const double* __restrict__ A, long lda, int K1, const double* __restrict__ Bm, long ldb, int K2, int p,
                    double* __restrict__ Cpart) {
  extern __shared__ double sm[];
  double* sA = sm;                                  // K1 x (GRAM_TILE+1)
  double* sB = sm + static_cast<size_t>(K1) * (GRAM_TILE + 1);  // K2 x (GRAM_TILE+1)
  const int f_begin = blockIdx.x * GRAM_CHUNK, f_end = min(p, f_begin + GRAM_CHUNK);
  const int nout = K1 * K2;
  // each thread owns outputs o = tid, tid+256, ... (at most 16 for K1=K2=64)
  double acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.0;
  for (int f0 = f_begin; f0 < f_end; f0 += GRAM_TILE) {
    const int nf = min(GRAM_TILE, f_end - f0);
    __syncthreads();
    for (int e = threadIdx.x; e < (K1 + K2) * GRAM_TILE; e += blockDim.x) {
      const int row = e / GRAM_TILE, f = e % GRAM_TILE;
      double v = 0.0;
      if (f < nf) v = row < K1 ? A[static_cast<size_t>(row) * lda + f0 + f] : Bm[static_cast<size_t>(row - K1) * ldb + f0 + f];
      if (row < K1) sA[row * (GRAM_TILE + 1) + f] = v;
      else sB[(row - K1) * (GRAM_TILE + 1) + f] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int o = threadIdx.x + k * 256;
      if (o < nout) {
        const int i = o / K2, j = o % K2;
        double s = acc[k];
#pragma unroll 8
        for (int f = 0; f < GRAM_TILE; ++f) s = fma(sA[i * (GRAM_TILE + 1) + f], sB[j * (GRAM_TILE + 1) + f], s);
        acc[k] = s;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int o = threadIdx.x + k * 256;
    if (o < nout) Cpart[static_cast<size_t>(blockIdx.x) * nout + o] = acc[k];
  }
}

// Few outputs (K1 * K2 <= 16: the q x q, K x 1 and 1 x 1 contractions of SIMPLS / UNIPALS / KERNEL, most of them with q = 1):
// the tiled kernel above spends its time in 64 shared-memory staging rounds per chunk (75 us per call at p = 50,000, 48 calls per
// SIMPLS fit at C2).  Here a thread takes every 256th feature of the chunk with all outputs in registers (coalesced loads),
// then one fixed-order block sum per output.  Same Cpart layout, so reduce_chunks applies unchanged.
template <int NOUT>
__global__ void __launch_bounds__(256)
gram_small_kernel(const double* __restrict__ A, long lda, int K1, const double* __restrict__ Bm, long ldb, int K2, int p,
                  double* __restrict__ Cpart) {
  __shared__ double scratch[32 * NOUT];
  const int f_begin = blockIdx.x * GRAM_CHUNK, f_end = min(p, f_begin + GRAM_CHUNK);
  const int nout = K1 * K2;
  double acc[NOUT];
#pragma unroll
  for (int o = 0; o < NOUT; ++o) acc[o] = 0.0;
  for (int f = f_begin + threadIdx.x; f < f_end; f += 256) {
#pragma unroll
    for (int o = 0; o < NOUT; ++o) {
      if (o < nout) {
        const int i = o / K2, j = o - i * K2;
        acc[o] = fma(A[static_cast<size_t>(i) * lda + f], Bm[static_cast<size_t>(j) * ldb + f], acc[o]);
      }
    }
  }
  block_sum<NOUT>(acc, scratch);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int o = 0; o < NOUT; ++o)
      if (o < nout) Cpart[static_cast<size_t>(blockIdx.x) * nout + o] = acc[o];
  }
}

__global__ void __launch_bounds__(256)
reduce_chunks_kernel(const double* __restrict__ Cpart, int nchunks, int len, double* __restrict__ C) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= len) return;
  double s = 0.0;
  for (int c = 0; c < nchunks; ++c) s += Cpart[static_cast<size_t>(c) * len + o];
  C[o] = s;
}

// out[c][j] = sum_k in[k][j] * rowscale[k] * M[k*C + c]
__global__ void __launch_bounds__(256)
right_multiply_kernel(const double* __restrict__ in, long ldin, int K, int p, const double* __restrict__ rowscale,
                      const double* __restrict__ M, int C, double* __restrict__ out, long ldout) {
  extern __shared__ double sM[];  // K x C, pre-scaled
  for (int e = threadIdx.x; e < K * C; e += blockDim.x) sM[e] = M[e] * (rowscale ? rowscale[e / C] : 1.0);
  __syncthreads();
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < p; j += gridDim.x * blockDim.x) {
    for (int c = 0; c < C; ++c) {
      double s = 0.0;
      for (int k = 0; k < K; ++k) s = fma(in[static_cast<size_t>(k) * ldin + j], sM[k * C + c], s);
      out[static_cast<size_t>(c) * ldout + j] = s;
    }
  }
}

// out_part[(s*C + c0 + c)*ldo + i] = sum_{j in split s} nan0(Xt[j][i]) * Bm[(c0+c)*ldb + j],  c < NC
// FB features per step = independent 16-byte loads in flight per thread: 8 for narrow outputs (ncu at the C5 predict shape,
// 1 M samples x 2,000 features, showed 79 % of the stall samples on the loads with 4 in flight), 4 when NC = 8 / 16
// accumulator pairs already fill the registers.
template <int NC, int FB>
__global__ void __launch_bounds__(256, (NC <= 4 ? 3 : 1))
skinny_gemm_kernel(const double* __restrict__ Xt, long ld, int n, const double* __restrict__ Bm, long ldb, int C, int c0,
                   const int* __restrict__ split_f0, const int* __restrict__ split_f1, double* __restrict__ out_part,
                   long ldo, const double* __restrict__ mean, const double* __restrict__ scale, int* __restrict__ flag) {
  const int s = blockIdx.y;
  const int f0 = split_f0[s], f1 = split_f1[s];
  const int r = (blockIdx.x * 256 + threadIdx.x) * 2;
  if (r >= n) return;
  const int nc = min(NC, C - c0);
  const bool last_is_pad = (r + 1 >= n);  // odd n: the second lane element is the zero padding, not a sample
  int bad = 0;
  double ax[NC], ay[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) ax[c] = ay[c] = 0.0;
  const double* __restrict__ xp = Xt + r;
  int j = f0;
  for (; j + FB <= f1; j += FB) {
    double2 x[FB];
#pragma unroll
    for (int k = 0; k < FB; ++k) {
      x[k] = ld_stream(reinterpret_cast<const double2*>(xp + static_cast<size_t>(j + k) * ld));
      if (last_is_pad) x[k].y = 0.0;  // element n of an odd-length feature is not a sample (an adopted view may hold anything there)
      bad |= !isfinite(x[k].x) | !isfinite(x[k].y);
      if (mean) {  // StandardScaler.transform fused into the product (mbpls.py:1097,:1369): z = (x - mean) / scale
        const double m = __ldg(mean + j + k);
        x[k].x -= m;
        x[k].y -= m;
        if (scale) {  // (scale == NULL: the caller folded 1 / scale into Bm, no fp64 division in the streaming loop)
          const double sc = __ldg(scale + j + k);
          x[k].x /= sc;
          x[k].y /= sc;
        }
      }
      if (isnan(x[k].x)) x[k].x = 0.0;
      if (isnan(x[k].y)) x[k].y = 0.0;
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c < nc) {
#pragma unroll
        for (int k = 0; k < FB; ++k) {
          const double bv = __ldg(Bm + static_cast<size_t>(c0 + c) * ldb + j + k);
          ax[c] = fma(x[k].x, bv, ax[c]);
          ay[c] = fma(x[k].y, bv, ay[c]);
        }
      }
    }
  }
  for (; j < f1; ++j) {
    double2 x = ld_stream(reinterpret_cast<const double2*>(xp + static_cast<size_t>(j) * ld));
    if (last_is_pad) x.y = 0.0;
    bad |= !isfinite(x.x) | !isfinite(x.y);
    if (mean) {
      const double m = __ldg(mean + j);
      x.x -= m;
      x.y -= m;
      if (scale) {
        const double sc = __ldg(scale + j);
        x.x /= sc;
        x.y /= sc;
      }
    }
    if (isnan(x.x)) x.x = 0.0;
    if (isnan(x.y)) x.y = 0.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (c < nc) {
        const double bv = __ldg(Bm + static_cast<size_t>(c0 + c) * ldb + j);
        ax[c] = fma(x.x, bv, ax[c]);
        ay[c] = fma(x.y, bv, ay[c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    if (c < nc) {
      double* o = out_part + (static_cast<size_t>(s) * C + c0 + c) * ldo + r;
      *reinterpret_cast<double2*>(o) = make_double2(ax[c], last_is_pad ? 0.0 : ay[c]);
    }
  }
  if (flag && bad) *flag = 1;  // idempotent: lets the caller reject NaN/inf input without an extra pass
}

// ------------------------------------------------------------------------------------------
// Tall batches (predict / superscores on m >> p new samples, BASELINE config 5): the same product as skinny_gemm_kernel
// fed by a shared-memory ring instead of per-thread loads.  With one feature 8 MB long every thread-level load of the
// kernel above touches a different row; ncu (1 M x 2,000 -> 4) showed 79 % of its stall samples on those loads with
// <= 74 KB in flight per SM (3.3 TB/s).  Here a persistent CTA owns tiles of TS samples: a producer warp streams the
// tile's chunk of every feature (TS * 8 bytes, 1-D bulk / TMA copy) through a ring of TR stages tracked by full / empty
// mbarriers, 8 consumer warps hold the tile's NC x TS outputs in registers.  One CTA covers all features of its tile, so
// results are written directly (no split partials, fixed summation order).  TWO CTAs per SM on half rings (2 x 6 stages x
// 16 KB = 192 KB in flight per SM): with one CTA the issue slots were 53 % busy behind `wait` / `math_pipe_throttle` stalls of
// 2 warps per scheduler (ncu); a second CTA's warps fill them: 3.16 -> 2.71 ms for 16 GB = 5.9 TB/s = 0.90 of the copy peak.
// (16 consumer warps in ONE CTA, 4 samples each, were slower -- 3.56 ms: twice the barrier waits and coefficient broadcasts
// per chunk.)
// ------------------------------------------------------------------------------------------
#define TALL_TS 2048  // samples per tile = 16 KB per feature chunk
#define TALL_TR 6     // ring stages (96 KB per CTA, two CTAs per SM)
#define TALL_U 4      // 16-byte units per consumer thread (256 consumer threads x 4 x 2 samples = TS)
#define TALL_NC 256   // consumer threads + one producer warp

template <int NC>
__global__ void __launch_bounds__(TALL_NC + 32, 2)
skinny_tall_kernel(const double* __restrict__ Xt, long ld, int n, int p, const double* __restrict__ coef, int C,
                   double* __restrict__ out, long ldo, int* __restrict__ flag, int tile_len) {
  // coef: p x 8 doubles, row j = {mean_j (0 without centring), b_0j, b_1j, b_2j, b_3j, -, -, -}: travels through the ring with
  // the chunk of feature j, so the consumers read their coefficients as shared-memory broadcasts
  extern __shared__ __align__(128) unsigned char tall_smem[];
  double* ring = reinterpret_cast<double*>(tall_smem);
  double* cring = ring + static_cast<size_t>(TALL_TR) * TALL_TS;  // [TALL_TR][8]
  uint64_t* full = reinterpret_cast<uint64_t*>(cring + TALL_TR * 8);
  uint64_t* empty = full + TALL_TR;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // tile_len (a multiple of 16, <= TALL_TS) is chosen by the host so that the tiles divide evenly among the CTAs
  const int ntiles = (n + tile_len - 1) / tile_len;
  const int nmine = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long total = static_cast<long>(nmine) * p;  // chunks this CTA streams, in (tile, feature) order
  if (tid == 0) {
    for (int s = 0; s < TALL_TR; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], TALL_NC / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == TALL_NC / 32) {  // producer: one elected lane issues every bulk copy
    if (lane == 0) {
      for (long c = 0; c < total; ++c) {
        const int s = static_cast<int>(c % TALL_TR);
        if (c >= TALL_TR) mbar_wait(&empty[s], static_cast<uint32_t>(((c / TALL_TR) - 1) & 1));
        const int tile = blockIdx.x + static_cast<int>(c / p) * gridDim.x, j = static_cast<int>(c % p);
        const long i0 = static_cast<long>(tile) * tile_len;
        const uint32_t bytes = static_cast<uint32_t>(min(static_cast<long>(tile_len), ld - i0)) * 8u;  // ld % 16 == 0: whole 128-byte lines
        mbar_arrive_expect_tx(&full[s], bytes + 64u);
        bulk_g2s(ring + static_cast<size_t>(s) * TALL_TS, Xt + static_cast<size_t>(j) * ld + i0, bytes, &full[s]);
        bulk_g2s(cring + s * 8, coef + static_cast<size_t>(j) * 8, 64u, &full[s]);
      }
    }
    return;
  }

  const int nc = min(NC, C);
  int bad = 0;
  long c = 0;
  for (int t = 0; t < nmine; ++t) {
    const int tile = blockIdx.x + t * gridDim.x;
    const long i0 = static_cast<long>(tile) * tile_len;
    const int valid_units = static_cast<int>(min(static_cast<long>(tile_len), ld - i0) >> 1);
    // samples of this thread that lie beyond n (padding, or whatever an adopted view holds there) are not data
    bool live_x[TALL_U], live_y[TALL_U];
#pragma unroll
    for (int k = 0; k < TALL_U; ++k) {
      const long i = i0 + 2 * (tid + k * TALL_NC);
      live_x[k] = i < n && (tid + k * TALL_NC) < valid_units;
      live_y[k] = i + 1 < n && (tid + k * TALL_NC) < valid_units;
    }
    double2 acc[TALL_U][NC];
#pragma unroll
    for (int k = 0; k < TALL_U; ++k)
#pragma unroll
      for (int cc = 0; cc < NC; ++cc) acc[k][cc] = make_double2(0.0, 0.0);
    for (int j = 0; j < p; ++j, ++c) {
      const int s = static_cast<int>(c % TALL_TR);
      mbar_wait(&full[s], static_cast<uint32_t>((c / TALL_TR) & 1));
      const double2* __restrict__ xs = reinterpret_cast<const double2*>(ring + static_cast<size_t>(s) * TALL_TS);
      const double2* __restrict__ cf = reinterpret_cast<const double2*>(cring + s * 8);
      double2 x[TALL_U];
#pragma unroll
      for (int k = 0; k < TALL_U; ++k) x[k] = live_x[k] ? xs[tid + k * TALL_NC] : make_double2(0.0, 0.0);
      const double2 c01 = cf[0], c23 = cf[1], c4 = cf[2];  // {mean, b0}, {b1, b2}, {b3, -}
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);  // the chunk sits in registers: hand the stage back
      const double m = c01.x;
      const double bv[4] = {c01.y, c23.x, c23.y, c4.x};
#pragma unroll
      for (int k = 0; k < TALL_U; ++k) {
        const int hx = __double2hiint(x[k].x), hy = __double2hiint(x[k].y);
        double zx = x[k].x - m, zy = x[k].y - m;  // StandardScaler.transform: 1 / scale is folded into the coefficients
        if (((hx & 0x7ff00000) == 0x7ff00000) | ((hy & 0x7ff00000) == 0x7ff00000)) {  // rare: NaN or +-inf in this unit
          if (live_x[k] && ((hx & 0x7ff00000) == 0x7ff00000)) bad = 1;
          if (live_y[k] && ((hy & 0x7ff00000) == 0x7ff00000)) bad = 1;
          if (isnan(zx)) zx = 0.0;  // NaN entries count as zero after scaling (:1379-1383)
          if (isnan(zy)) zy = 0.0;
        }
        if (!live_x[k]) zx = 0.0;
        if (!live_y[k]) zy = 0.0;
#pragma unroll
        for (int cc = 0; cc < NC; ++cc) {
          acc[k][cc].x = fma(zx, bv[cc], acc[k][cc].x);
          acc[k][cc].y = fma(zy, bv[cc], acc[k][cc].y);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < TALL_U; ++k) {
      const long i = i0 + 2 * (tid + k * TALL_NC);
      if (live_x[k]) {
#pragma unroll
        for (int cc = 0; cc < NC; ++cc)
          if (cc < nc) {
            double* o = out + static_cast<size_t>(cc) * ldo + i;
            *reinterpret_cast<double2*>(o) = make_double2(acc[k][cc].x, live_y[k] ? acc[k][cc].y : 0.0);
          }
      }
    }
  }
  if (flag && bad) *flag = 1;
}

__global__ void __launch_bounds__(256)
rank1_update_kernel(double* __restrict__ Xt, long ld, int n, int p, const double* __restrict__ ts,
                    const double* __restrict__ pvec) {
  const int j = blockIdx.y;
  if (j >= p) return;
  const double pj = pvec[j];
  double* x = Xt + static_cast<size_t>(j) * ld;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    x[i] = __dsub_rn(x[i], __dmul_rn(ts[i], pj));
}

// one CTA per row; four independent loads in flight per thread (a row can be 8 MB: K x p weights at the headline size)
__global__ void __launch_bounds__(1024) rows_sumsq_kernel(const double* __restrict__ M, long ld, int n, double* __restrict__ out) {
  __shared__ double scratch[32];
  const double* m = M + static_cast<size_t>(blockIdx.x) * ld;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  const int nt = blockDim.x;
  int i = threadIdx.x;
  for (; i + 3 * nt < n; i += 4 * nt) {
    const double a0 = m[i], a1 = m[i + nt], a2 = m[i + 2 * nt], a3 = m[i + 3 * nt];
    s0 = fma(a0, a0, s0);
    s1 = fma(a1, a1, s1);
    s2 = fma(a2, a2, s2);
    s3 = fma(a3, a3, s3);
  }
  for (; i < n; i += nt) s0 = fma(m[i], m[i], s0);
  const double s = block_sum1((s0 + s1) + (s2 + s3), scratch);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256)
rows_scale_kernel(double* __restrict__ M, long ld, int n, const double* __restrict__ scale, int divide) {
  double* m = M + static_cast<size_t>(blockIdx.y) * ld;
  const double s = scale[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m[i] = divide ? m[i] / s : m[i] * s;
}

extern "C" {

int mbpls_abi_version(void) { return MBPLS_ABI_VERSION; }

int mbpls_gram_num_chunks(int p) { return p <= 0 ? 0 : (p + GRAM_CHUNK - 1) / GRAM_CHUNK; }

int mbpls_gram_partial_f64(const double* A, long lda, int K1, const double* Bm, long ldb, int K2, int p, double* Cpart,
                           void* stream) {
  if (!A || !Bm || !Cpart || K1 < 1 || K2 < 1) return MBPLS_ERR_ARG;
  if (K1 > 64 || K2 > 64) return MBPLS_ERR_SIZE;
  if (p <= 0) return MBPLS_OK;
  const int grid = mbpls_gram_num_chunks(p);
  const int nout = K1 * K2;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (nout == 1) {
    gram_small_kernel<1><<<grid, 256, 0, st>>>(A, lda, K1, Bm, ldb, K2, p, Cpart);
    MBPLS_RETURN_LAST();
  }
  if (nout <= 4) {
    gram_small_kernel<4><<<grid, 256, 0, st>>>(A, lda, K1, Bm, ldb, K2, p, Cpart);
    MBPLS_RETURN_LAST();
  }
  if (nout <= 16) {
    gram_small_kernel<16><<<grid, 256, 0, st>>>(A, lda, K1, Bm, ldb, K2, p, Cpart);
    MBPLS_RETURN_LAST();
  }
  if (nout <= 64 && (K1 == 1 || K2 == 1)) {  // up to 64 earlier components against one vector
    gram_small_kernel<64><<<grid, 256, 0, st>>>(A, lda, K1, Bm, ldb, K2, p, Cpart);
    MBPLS_RETURN_LAST();
  }
  const size_t smem = static_cast<size_t>(K1 + K2) * (GRAM_TILE + 1) * sizeof(double);
  gram_partial_kernel<<<grid, 256, smem, st>>>(A, lda, K1, Bm, ldb, K2, p, Cpart);
  MBPLS_RETURN_LAST();
}

int mbpls_reduce_chunks_f64(const double* Cpart, int nchunks, int len, double* C, void* stream) {
  if (!Cpart || !C || nchunks < 0 || len < 0) return MBPLS_ERR_ARG;
  if (len == 0) return MBPLS_OK;
  reduce_chunks_kernel<<<(len + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(Cpart, nchunks, len, C);
  MBPLS_RETURN_LAST();
}

int mbpls_right_multiply_f64(const double* in, long ldin, int K, int p, const double* rowscale, const double* M, int C,
                             double* out, long ldout, void* stream) {
  if (!in || !M || !out || K < 1 || C < 1) return MBPLS_ERR_ARG;
  if (static_cast<size_t>(K) * C * sizeof(double) > 48 * 1024) return MBPLS_ERR_SIZE;
  if (p <= 0) return MBPLS_OK;
  int grid = (p + 255) / 256;
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  right_multiply_kernel<<<grid, 256, static_cast<size_t>(K) * C * sizeof(double), static_cast<cudaStream_t>(stream)>>>(
      in, ldin, K, p, rowscale, M, C, out, ldout);
  MBPLS_RETURN_LAST();
}

int mbpls_skinny_gemm_f64(const double* Xt, long ld, int n, const double* Bm, long ldb, int C, const int* split_f0,
                          const int* split_f1, int nsplit, double* out_part, long ldo, const double* mean,
                          const double* scale, int* nonfinite_flag, void* stream) {
  if (scale != nullptr && mean == nullptr) return MBPLS_ERR_ARG;  // mean alone: centring only (1 / scale folded into Bm)
  if (!Xt || !Bm || !split_f0 || !split_f1 || !out_part || C < 1 || (ld % 2) != 0 || (ldo % 2) != 0) return MBPLS_ERR_ARG;
  if (nsplit == 0 || n == 0) return MBPLS_OK;
  if (nsplit > 65535) return MBPLS_ERR_SIZE;
  dim3 grid((n + 511) / 512, nsplit);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int c0 = 0; c0 < C; c0 += 16) {
    const int nc = C - c0;
    if (nc >= 16 || nc > 8) skinny_gemm_kernel<16, 4><<<grid, 256, 0, st>>>(Xt, ld, n, Bm, ldb, C, c0, split_f0, split_f1, out_part, ldo, mean, scale, nonfinite_flag);
    else if (nc > 4) skinny_gemm_kernel<8, 4><<<grid, 256, 0, st>>>(Xt, ld, n, Bm, ldb, C, c0, split_f0, split_f1, out_part, ldo, mean, scale, nonfinite_flag);
    else if (nc > 2) skinny_gemm_kernel<4, 6><<<grid, 256, 0, st>>>(Xt, ld, n, Bm, ldb, C, c0, split_f0, split_f1, out_part, ldo, mean, scale, nonfinite_flag);
    else if (nc > 1) skinny_gemm_kernel<2, 8><<<grid, 256, 0, st>>>(Xt, ld, n, Bm, ldb, C, c0, split_f0, split_f1, out_part, ldo, mean, scale, nonfinite_flag);
    else skinny_gemm_kernel<1, 8><<<grid, 256, 0, st>>>(Xt, ld, n, Bm, ldb, C, c0, split_f0, split_f1, out_part, ldo, mean, scale, nonfinite_flag);
  }
  MBPLS_RETURN_LAST();
}

/* tall batches: same product for C <= 4 outputs and ALL p features, written directly to out[c*ldo + i] (no partials).
 * Persistent CTAs stream 16 KB chunks of every feature through a 192 KB TMA ring (csrc/finalize.cu skinny_tall_kernel).
 * coef: p x 8 doubles, row j = {mean_j or 0, b_0j .. b_3j, 0, 0, 0} with 1 / scale_j already folded into the b's. */
int mbpls_skinny_gemm_tall_f64(const double* Xt, long ld, int n, int p, const double* coef, int C, double* out, long ldo,
                               int* nonfinite_flag, void* stream) {
  if (!Xt || !coef || !out || C < 1 || C > 4 || (ld % 16) != 0 || (ldo % 2) != 0 || ld < n) return MBPLS_ERR_ARG;
  if (n == 0 || p == 0) return MBPLS_OK;
  const size_t smem = static_cast<size_t>(TALL_TR) * TALL_TS * sizeof(double) + TALL_TR * 64 + 2 * TALL_TR * sizeof(uint64_t) + 64;
  if (smem > static_cast<size_t>(smem_optin())) return MBPLS_ERR_SIZE;
  // Tiles of TALL_TS samples leave a ragged last round (1 M samples: 489 tiles on 148 CTAs = 3.3 rounds, i.e. 4 for some and 3
  // for the rest: 0.82 efficiency, measured 5.1 TB/s).  Cut the sample axis into a whole number of rounds instead: the smallest
  // k with ceil(n / (k * CTAs)) <= TALL_TS, tiles of that length rounded up to 16 samples.
  int tile_len = TALL_TS;
  {
    const long sms = 2L * num_sms();  // two CTAs per SM
    const long k = (static_cast<long>(n) + sms * TALL_TS - 1) / (sms * TALL_TS);
    long t = (static_cast<long>(n) + k * sms - 1) / (k * sms);
    t = (t + 15) / 16 * 16;
    if (t < 16) t = 16;
    if (t < TALL_TS) tile_len = static_cast<int>(t);
  }
  const int ntiles = (n + tile_len - 1) / tile_len;
  const int grid = ntiles < 2 * num_sms() ? ntiles : 2 * num_sms();
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define TALL_LAUNCH(NCV)                                                                                                   \
  do {                                                                                                                     \
    cudaFuncSetAttribute(skinny_tall_kernel<NCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));    \
    skinny_tall_kernel<NCV><<<grid, TALL_NC + 32, smem, st>>>(Xt, ld, n, p, coef, C, out, ldo, nonfinite_flag, tile_len);           \
  } while (0)
  if (C > 2) TALL_LAUNCH(4);
  else if (C > 1) TALL_LAUNCH(2);
  else TALL_LAUNCH(1);
#undef TALL_LAUNCH
  MBPLS_RETURN_LAST();
}

int mbpls_rank1_update_f64(double* Xt, long ld, int n, int p, const double* ts, const double* pvec, void* stream) {
  if (!Xt || !ts || !pvec) return MBPLS_ERR_ARG;
  if (p == 0 || n == 0) return MBPLS_OK;
  int gx = (n + 255) / 256;
  if (gx > 64) gx = 64;
  for (int j0 = 0; j0 < p; j0 += 65535) {
    const int pj = (p - j0) < 65535 ? (p - j0) : 65535;
    dim3 grid(gx, pj);
    rank1_update_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(Xt + static_cast<size_t>(j0) * ld, ld, n, pj, ts,
                                                                            pvec + j0);
  }
  MBPLS_RETURN_LAST();
}

int mbpls_rows_sumsq_f64(const double* M, long ld, int rows, int n, double* out, void* stream) {
  if (!M || !out || rows < 0) return MBPLS_ERR_ARG;
  if (rows == 0) return MBPLS_OK;
  rows_sumsq_kernel<<<rows, 1024, 0, static_cast<cudaStream_t>(stream)>>>(M, ld, n, out);
  MBPLS_RETURN_LAST();
}

int mbpls_rows_scale_f64(double* M, long ld, int rows, int n, const double* scale, int divide, void* stream) {
  if (!M || !scale || rows < 0 || rows > 65535) return MBPLS_ERR_ARG;
  if (rows == 0 || n == 0) return MBPLS_OK;
  int gx = (n + 255) / 256;
  if (gx > 256) gx = 256;
  dim3 grid(gx, rows);
  rows_scale_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(M, ld, n, scale, divide);
  MBPLS_RETURN_LAST();
}

}  // extern "C"
