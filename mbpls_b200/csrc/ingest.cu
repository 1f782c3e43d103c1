// Ingest + preprocessing kernels: row-major -> feature-major transpose, NaN census, fused
// column standardisation (the reference's StandardScaler.fit_transform, mbpls/mbpls.py:307,314,
// 325-326), standardise-apply for new data (:1097,:1369), per-feature sums of squares
// (varx / varxblocks, :822-830, :938-945).
#include "stream.cuh"
#include "launch.cuh"

using namespace mbpls;

// ------------------------------------------------------------------------------------------
// transpose: src is a row-major chunk (rows x cols, leading dim lds) of one block; it lands in
// the feature-major matrix at features [0, cols) (dst already offset to the block) and samples
// [row0, row0 + rows).
// ------------------------------------------------------------------------------------------
template <class T>  // T = double, or float: single-precision sources are widened here, on the device, after a half-size upload
__global__ void __launch_bounds__(256) transpose_in_kernel(const T* __restrict__ src, long lds, int rows, int cols,
                                                           double* __restrict__ dst, long ld, int row0) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + 8 * k][tx] = static_cast<double>(src[static_cast<size_t>(r) * lds + c]);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, r = r0 + tx;
    if (r < rows && c < cols) dst[static_cast<size_t>(c) * ld + row0 + r] = tile[tx][ty + 8 * k];
  }
}

// feature-major -> row-major (used to hand standardised blocks back when copy=False and for tests)
__global__ void __launch_bounds__(256) transpose_out_kernel(const double* __restrict__ src, long ld, int rows, int cols,
                                                            double* __restrict__ dst, long ldd, int row0) {
  __shared__ double tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + ty + 8 * k, r = r0 + tx;
    if (r < rows && c < cols) tile[ty + 8 * k][tx] = src[static_cast<size_t>(c) * ld + row0 + r];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = r0 + ty + 8 * k, c = c0 + tx;
    if (r < rows && c < cols) dst[static_cast<size_t>(r) * ldd + c] = tile[tx][ty + 8 * k];
  }
}

// ------------------------------------------------------------------------------------------
// NaN census (mbpls/mbpls.py:255-271): per-feature NaN count and per-(block,row) NaN flag.
// One warp per feature; row flags are idempotent byte stores (every writer stores 1).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) census_kernel(const double* __restrict__ Xt, long ld, int n, int p,
                                                     const int* __restrict__ block_off, int B,
                                                     int* __restrict__ col_nan, unsigned char* __restrict__ row_flag,
                                                     long ldf, int* __restrict__ inf_flag) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < p; j += warps) {
    const int b = block_of(block_off, B, j);
    const double* x = Xt + static_cast<size_t>(j) * ld;
    int cnt = 0, inf = 0;
    for (int i = lane; i < n; i += 32) {
      const double v = x[i];
      if (isnan(v)) {
        ++cnt;
        if (row_flag) row_flag[static_cast<size_t>(b) * ldf + i] = 1;
      }
      inf |= isinf(v);
    }
    if (inf && inf_flag) *inf_flag = 1;  // check_array rejects infinities even when NaN is allowed
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(MBPLS_FULL_MASK, cnt, o);
    if (lane == 0) col_nan[j] = cnt;
  }
}

// ------------------------------------------------------------------------------------------
// Standardisation arithmetic shared by the fused (smem-resident) and the fallback kernels.
// scikit-learn semantics: nan-aware mean, corrected two-pass population variance, near-constant
// features get scale 1 (sklearn/preprocessing/_data.py::_is_constant_feature).
// ------------------------------------------------------------------------------------------
struct ScalerOut {
  double* mean;
  double* var;
  double* scale;
  long long* seen;  // observed (non-NaN) count per feature == n_samples_seen_
  double* zss;      // nansum of the standardised feature squared (for varx)
};

__device__ __forceinline__ void finish_stats(double cnt, double sum, double& mean) { mean = sum / cnt; }

__device__ __forceinline__ double scale_from(double cnt, double mean, double corr, double ssq, double& var) {
  ssq -= corr * corr / cnt;
  var = ssq / cnt;
  const double eps = 2.220446049250313e-16;
  const double bound = cnt * eps * var + (cnt * mean * eps) * (cnt * mean * eps);
  return (var <= bound) ? 1.0 : sqrt(var);
}

// T threads (a warp, or the whole CTA) cooperate on one resident feature.
template <bool CTA_WIDE>
__device__ __forceinline__ void standardize_resident(double* x, int n, int j, const ScalerOut& o, double* scratch) {
  const int tid = CTA_WIDE ? threadIdx.x : (threadIdx.x & 31);
  const int nt = CTA_WIDE ? blockDim.x : 32;
  double v[2] = {0.0, 0.0};  // count, sum
  for (int i = tid; i < n; i += nt) {
    const double xi = x[i];
    if (!isnan(xi)) {
      v[0] += 1.0;
      v[1] += xi;
    }
  }
  if (CTA_WIDE) block_sum<2>(v, scratch);
  else { v[0] = warp_sum(v[0]); v[1] = warp_sum(v[1]); }
  const double cnt = v[0], mean = v[1] / v[0];
  double c[2] = {0.0, 0.0};  // correction, sum of squares
  for (int i = tid; i < n; i += nt) {
    const double d = x[i] - mean;
    if (!isnan(d)) {
      c[0] += d;
      c[1] += d * d;
    }
  }
  if (CTA_WIDE) block_sum<2>(c, scratch);
  else { c[0] = warp_sum(c[0]); c[1] = warp_sum(c[1]); }
  double var;
  const double scale = scale_from(cnt, mean, c[0], c[1], var);
  double z[1] = {0.0};
  const UniformDivisor by_scale(scale);  // bit-identical to `/ scale` (StandardScaler's true division), a tenth of the instructions
  for (int i = tid; i < n; i += nt) {
    double zi = x[i] - mean;
    zi = by_scale(zi);
    x[i] = zi;
    if (!isnan(zi)) z[0] += zi * zi;
  }
  if (CTA_WIDE) block_sum<1>(z, scratch);
  else z[0] = warp_sum(z[0]);
  if (tid == 0) {
    o.mean[j] = mean;
    o.var[j] = var;
    o.scale[j] = scale;
    o.seen[j] = static_cast<long long>(cnt);
    o.zss[j] = z[0];
  }
}

// Short features: one consumer warp per resident feature (8 consumer warps + the producer warp).
struct StandardizeWarpOp {
  int n;
  long ld;
  ScalerOut out;
  __device__ __forceinline__ void operator()(double* slab, int f0, int nf) {
    const int warp = threadIdx.x >> 5, nw = (blockDim.x >> 5) - 1;
    for (int f = warp; f < nf; f += nw) standardize_resident<false>(slab + static_cast<size_t>(f) * ld, n, f0 + f, out, nullptr);
  }
};

__global__ void __launch_bounds__(288) standardize_warp_kernel(double* __restrict__ Xt, StreamShape sh, int n, ScalerOut out) {
  StandardizeWarpOp op{n, sh.ld, out};
  stream_feature_slabs<true>(Xt, sh, op);
}

// Long features: 1024-thread CTA (31 consumer warps + the producer warp); each consumer keeps its EPT
// samples of the resident feature in registers across the three reductions (one shared-memory read
// and one write per element).
template <int EPT>
struct StandardizeWideOp {
  int n;
  long ld;
  ScalerOut out;
  double* scratch;
  __device__ __forceinline__ void operator()(double* slab, int f0, int nf) {
    const int nt = blockDim.x - 32;
    for (int f = 0; f < nf; ++f) {
      double* x = slab + static_cast<size_t>(f) * ld;
      double xr[EPT];
      double v[2] = {0.0, 0.0};
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const int i = threadIdx.x + k * nt;
        xr[k] = i < n ? x[i] : NAN;  // out-of-range lanes behave like missing values
        if (!isnan(xr[k])) {
          v[0] += 1.0;
          v[1] += xr[k];
        }
      }
      block_sum_consumers<2>(v, scratch, nt);
      const double cnt = v[0], mean = v[1] / v[0];
      double c[2] = {0.0, 0.0};
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        xr[k] = xr[k] - mean;
        if (!isnan(xr[k])) {
          c[0] += xr[k];
          c[1] = fma(xr[k], xr[k], c[1]);
        }
      }
      block_sum_consumers<2>(c, scratch, nt);
      double var;
      const double scale = scale_from(cnt, mean, c[0], c[1], var);
      double z[1] = {0.0};
      const UniformDivisor by_scale(scale);
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const int i = threadIdx.x + k * nt;
        const double zi = by_scale(xr[k]);
        if (i < n) x[i] = zi;
        if (!isnan(zi)) z[0] = fma(zi, zi, z[0]);
      }
      block_sum_consumers<1>(z, scratch, nt);
      if (threadIdx.x == 0) {
        out.mean[f0 + f] = mean;
        out.var[f0 + f] = var;
        out.scale[f0 + f] = scale;
        out.seen[f0 + f] = static_cast<long long>(cnt);
        out.zss[f0 + f] = z[0];
      }
    }
  }
};

template <int EPT>
__global__ void __launch_bounds__(1024, 1) standardize_wide_kernel(double* __restrict__ Xt, StreamShape sh, int n, ScalerOut out) {
  __shared__ double scratch[64];
  StandardizeWideOp<EPT> op{n, sh.ld, out, scratch};
  stream_feature_slabs<true>(Xt, sh, op);
}

template <int EPT>
static void launch_standardize_wide(double* Xt, const StreamShape& sh, int n, const ScalerOut& out, int grid, size_t smem,
                                    cudaStream_t st) {
  cudaFuncSetAttribute(standardize_wide_kernel<EPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  standardize_wide_kernel<EPT><<<grid, 1024, smem, st>>>(Xt, sh, n, out);
}

// Mid-length features (1024 < n <= 16384): register-resident variant (see loadings_deflate_regs_kernel):
// a 256-thread CTA holds one whole feature in registers across the three reductions.
template <int EPT2>
__global__ void __launch_bounds__(256, 2) standardize_regs_kernel(double* __restrict__ Xt, long ld, int n, int p, ScalerOut out) {
  __shared__ double scratch[64];
  const int n2 = (n + 1) >> 1;
  const bool odd = (n & 1) != 0;
  double2 xr[EPT2];
  int j = blockIdx.x;
  if (j < p) {
    const double2* x2 = reinterpret_cast<const double2*>(Xt + static_cast<size_t>(j) * ld);
#pragma unroll
    for (int k = 0; k < EPT2; ++k) {
      const int i = threadIdx.x + k * 256;
      xr[k] = i < n2 ? ld_stream(x2 + i) : make_double2(NAN, NAN);  // out of range == missing
    }
  }
  while (j < p) {
    double v[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < EPT2; ++k) {
      const int i = threadIdx.x + k * 256;
      if (odd && i == n2 - 1) xr[k].y = NAN;  // the padding element is not a sample
      if (!isnan(xr[k].x)) { v[0] += 1.0; v[1] += xr[k].x; }
      if (!isnan(xr[k].y)) { v[0] += 1.0; v[1] += xr[k].y; }
    }
    block_sum<2>(v, scratch);
    const double cnt = v[0], mean = v[1] / v[0];
    double c[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < EPT2; ++k) {
      xr[k].x -= mean;
      xr[k].y -= mean;
      if (!isnan(xr[k].x)) { c[0] += xr[k].x; c[1] = fma(xr[k].x, xr[k].x, c[1]); }
      if (!isnan(xr[k].y)) { c[0] += xr[k].y; c[1] = fma(xr[k].y, xr[k].y, c[1]); }
    }
    block_sum<2>(c, scratch);
    double var;
    const double scale = scale_from(cnt, mean, c[0], c[1], var);
    const int jn = j + gridDim.x;
    double2* x2 = reinterpret_cast<double2*>(Xt + static_cast<size_t>(j) * ld);
    const double2* xn = reinterpret_cast<const double2*>(Xt + static_cast<size_t>(jn < p ? jn : j) * ld);
    double z[1] = {0.0};
    const UniformDivisor by_scale(scale);
#pragma unroll
    for (int k = 0; k < EPT2; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < n2) {
        double2 zi = make_double2(by_scale(xr[k].x), by_scale(xr[k].y));
        if (!isnan(zi.x)) z[0] = fma(zi.x, zi.x, z[0]);
        if (!isnan(zi.y)) z[0] = fma(zi.y, zi.y, z[0]);
        if (odd && i == n2 - 1) zi.y = 0.0;  // keep the padding element zero
        st_stream(x2 + i, zi);
      }
    }
    if (jn < p) {  // registers are drained: the next feature's loads fly during the reduction below
#pragma unroll
      for (int k = 0; k < EPT2; ++k) {
        const int i = threadIdx.x + k * 256;
        xr[k] = i < n2 ? ld_stream(xn + i) : make_double2(NAN, NAN);
      }
    }
    block_sum<1>(z, scratch);
    if (threadIdx.x == 0) {
      out.mean[j] = mean;
      out.var[j] = var;
      out.scale[j] = scale;
      out.seen[j] = static_cast<long long>(cnt);
      out.zss[j] = z[0];
    }
    j = jn;
  }
}

template <int EPT2>
static void launch_standardize_regs(double* Xt, long ld, int n, int p, const ScalerOut& out, cudaStream_t st) {
  int grid = num_sms() * 2;
  if (grid > p) grid = p;
  standardize_regs_kernel<EPT2><<<grid, 256, 0, st>>>(Xt, ld, n, p, out);
}

// Fallback for features too long for shared memory: one CTA per feature straight from global
// memory (3 reads + 1 write instead of 1 + 1).
__global__ void __launch_bounds__(256) standardize_global_kernel(double* __restrict__ Xt, long ld, int n, int p, ScalerOut out) {
  __shared__ double scratch[64];
  for (int j = blockIdx.x; j < p; j += gridDim.x) {
    standardize_resident<true>(Xt + static_cast<size_t>(j) * ld, n, j, out, scratch);
    __syncthreads();
  }
}

// transform() of a fitted scaler on new data: z = (x - mean_j) / scale_j, in place, feature-major.
__global__ void __launch_bounds__(256) standardize_apply_kernel(double* __restrict__ Xt, long ld, int n, int p,
                                                                const double* __restrict__ mean,
                                                                const double* __restrict__ scale) {
  const int j = blockIdx.y;
  if (j >= p) return;
  const double m = mean[j];
  const UniformDivisor by_scale(scale[j]);
  double* x = Xt + static_cast<size_t>(j) * ld;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double z = x[i] - m;
    x[i] = by_scale(z);
  }
}

// inverse_transform for predictions: y = z * scale_c + mean_c (feature-major q x ld)
__global__ void __launch_bounds__(256) scaler_inverse_kernel(double* __restrict__ Zt, long ld, int n, int q,
                                                             const double* __restrict__ mean,
                                                             const double* __restrict__ scale) {
  const int c = blockIdx.y;
  if (c >= q) return;
  const double m = mean[c], s = scale[c];
  double* z = Zt + static_cast<size_t>(c) * ld;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double v = z[i] * s;
    z[i] = v + m;
  }
}

// nansum(x_j^2) per feature, one warp per feature.
__global__ void __launch_bounds__(256) feature_sumsq_kernel(const double* __restrict__ Xt, long ld, int n, int p,
                                                            double* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; j < p; j += warps) {
    const double* x = Xt + static_cast<size_t>(j) * ld;
    double s0 = 0.0, s1 = 0.0;
    int i = lane;
    for (; i + 32 < n; i += 64) {
      const double a = x[i], b = x[i + 32];
      if (!isnan(a)) s0 += a * a;
      if (!isnan(b)) s1 += b * b;
    }
    if (i < n) {
      const double a = x[i];
      if (!isnan(a)) s0 += a * a;
    }
    const double s = warp_sum(s0 + s1);
    if (lane == 0) out[j] = s;
  }
}

// Row-sharded standardisation (samples split over GPUs): the per-feature sums are all-reduced between passes, this
// finishes var_ / scale_ from the centred sums exactly like scale_from() does for the single-GPU kernels.
__global__ void __launch_bounds__(256)
scaler_finish_kernel(const double* __restrict__ corr, const double* __restrict__ ssq, const double* __restrict__ mean, double cnt,
                     double* __restrict__ var, double* __restrict__ scale, int p) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= p) return;
  double v;
  scale[j] = scale_from(cnt, mean[j], corr[j], ssq[j], v);
  var[j] = v;
}

// Deterministic segmented sum: out[s] = sum(v[off[s] .. off[s+1])), one 1024-thread CTA per segment, four independent
// loads in flight per thread (a block of 400k features used to take 0.7 ms with 256 threads and one load at a time).
__global__ void __launch_bounds__(1024) segsum_kernel(const double* __restrict__ v, const int* __restrict__ off,
                                                      double* __restrict__ out) {
  __shared__ double scratch[32];
  const int s = blockIdx.x;
  const int a = off[s], b = off[s + 1];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int i = a + threadIdx.x;
  for (; i + 3 * 1024 < b; i += 4 * 1024) {
    const double v0 = v[i], v1 = v[i + 1024], v2 = v[i + 2048], v3 = v[i + 3072];
    a0 += v0;
    a1 += v1;
    a2 += v2;
    a3 += v3;
  }
  for (; i < b; i += 1024) a0 += v[i];
  double acc = (a0 + a1) + (a2 + a3);
  acc = block_sum1(acc, scratch);
  if (threadIdx.x == 0) out[s] = acc;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int mbpls_transpose_in_f64(const double* src, long lds, int rows, int cols, double* dst, long ld, int row0, void* stream) {
  if (!src || !dst || rows < 0 || cols < 0) return MBPLS_ERR_ARG;
  if (rows == 0 || cols == 0) return MBPLS_OK;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  transpose_in_kernel<double><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, lds, rows, cols, dst, ld, row0);
  MBPLS_RETURN_LAST();
}

int mbpls_transpose_in_f32(const float* src, long lds, int rows, int cols, double* dst, long ld, int row0, void* stream) {
  if (!src || !dst || rows < 0 || cols < 0) return MBPLS_ERR_ARG;
  if (rows == 0 || cols == 0) return MBPLS_OK;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  transpose_in_kernel<float><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, lds, rows, cols, dst, ld, row0);
  MBPLS_RETURN_LAST();
}

int mbpls_transpose_out_f64(const double* src, long ld, int rows, int cols, double* dst, long ldd, int row0, void* stream) {
  if (!src || !dst || rows < 0 || cols < 0) return MBPLS_ERR_ARG;
  if (rows == 0 || cols == 0) return MBPLS_OK;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32);
  transpose_out_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(src, ld, rows, cols, dst, ldd, row0);
  MBPLS_RETURN_LAST();
}

int mbpls_nan_census_f64(const double* Xt, long ld, int n, int p, const int* block_off, int B, int* col_nan,
                         unsigned char* row_flag, long ldf, int* inf_flag, void* stream) {
  if (!Xt || !block_off || !col_nan || B < 1) return MBPLS_ERR_ARG;
  if (p == 0 || n == 0) return MBPLS_OK;
  int grid = (p + 7) / 8;
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  census_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(Xt, ld, n, p, block_off, B, col_nan, row_flag, ldf, inf_flag);
  MBPLS_RETURN_LAST();
}

// mode: 0 = auto (fused smem-resident pipeline when a feature fits, else global fallback),
//       1 = force the global-memory fallback.
int mbpls_standardize_fit_f64(double* Xt, long ld, int n, int p, double* mean, double* var, double* scale,
                              long long* seen, double* zss, int mode, void* stream) {
  if (!Xt || !mean || !var || !scale || !seen || !zss || ld < n || (ld % 16) != 0) return MBPLS_ERR_ARG;
  if (p == 0 || n == 0) return MBPLS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ScalerOut out{mean, var, scale, seen, zss};
  StreamShape sh;
  bool cta_wide = false;
  if (mode == 0 && n > 1024 && n <= 16384) {  // register-resident
    const int e = (((n + 1) >> 1) + 255) / 256;
    if (e <= 4) launch_standardize_regs<4>(Xt, ld, n, p, out, st);
    else if (e <= 8) launch_standardize_regs<8>(Xt, ld, n, p, out, st);
    else if (e <= 12) launch_standardize_regs<12>(Xt, ld, n, p, out, st);
    else if (e <= 16) launch_standardize_regs<16>(Xt, ld, n, p, out, st);
    else if (e <= 20) launch_standardize_regs<20>(Xt, ld, n, p, out, st);
    else if (e <= 24) launch_standardize_regs<24>(Xt, ld, n, p, out, st);
    else launch_standardize_regs<32>(Xt, ld, n, p, out, st);
  } else if ((mode == 0 || mode == 2) && pick_stream_shape(ld, p, &sh, &cta_wide)) {  // mode 2: force the smem pipeline
    const size_t smem = stream_smem_bytes(sh);
    const int grid = stream_grid(sh, smem);
    if (cta_wide) {
      const int ept = (n + 991) / 992;  // 31 consumer warps
      if (ept <= 4) launch_standardize_wide<4>(Xt, sh, n, out, grid, smem, st);
      else if (ept <= 8) launch_standardize_wide<8>(Xt, sh, n, out, grid, smem, st);
      else if (ept <= 12) launch_standardize_wide<12>(Xt, sh, n, out, grid, smem, st);
      else launch_standardize_wide<16>(Xt, sh, n, out, grid, smem, st);
    } else {
      cudaFuncSetAttribute(standardize_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      standardize_warp_kernel<<<grid, 288, smem, st>>>(Xt, sh, n, out);
    }
  } else {
    int grid = p < num_sms() * 8 ? p : num_sms() * 8;
    standardize_global_kernel<<<grid, 256, 0, st>>>(Xt, ld, n, p, out);
  }
  MBPLS_RETURN_LAST();
}

int mbpls_standardize_apply_f64(double* Xt, long ld, int n, int p, const double* mean, const double* scale, void* stream) {
  if (!Xt || !mean || !scale) return MBPLS_ERR_ARG;
  if (p == 0 || n == 0) return MBPLS_OK;
  int gx = (n + 255) / 256;
  if (gx > 64) gx = 64;
  for (int j0 = 0; j0 < p; j0 += 65535) {
    const int pj = (p - j0) < 65535 ? (p - j0) : 65535;
    dim3 grid(gx, pj);
    standardize_apply_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(Xt + static_cast<size_t>(j0) * ld, ld, n, pj,
                                                                                 mean + j0, scale + j0);
  }
  MBPLS_RETURN_LAST();
}

int mbpls_scaler_inverse_f64(double* Zt, long ld, int n, int q, const double* mean, const double* scale, void* stream) {
  if (!Zt || !mean || !scale || q > 65535) return MBPLS_ERR_ARG;
  if (q == 0 || n == 0) return MBPLS_OK;
  int gx = (n + 255) / 256;
  if (gx > 1024) gx = 1024;
  dim3 grid(gx, q);
  scaler_inverse_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(Zt, ld, n, q, mean, scale);
  MBPLS_RETURN_LAST();
}

int mbpls_feature_sumsq_f64(const double* Xt, long ld, int n, int p, double* out, void* stream) {
  if (!Xt || !out) return MBPLS_ERR_ARG;
  if (p == 0) return MBPLS_OK;
  int grid = (p + 7) / 8;
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  feature_sumsq_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(Xt, ld, n, p, out);
  MBPLS_RETURN_LAST();
}

int mbpls_scaler_finish_f64(const double* corr, const double* ssq, const double* mean, double count, double* var, double* scale,
                            int p, void* stream) {
  if (!corr || !ssq || !mean || !var || !scale || count <= 0) return MBPLS_ERR_ARG;
  if (p <= 0) return MBPLS_OK;
  scaler_finish_kernel<<<(p + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(corr, ssq, mean, count, var, scale, p);
  MBPLS_RETURN_LAST();
}

int mbpls_segsum_f64(const double* v, const int* off, int nseg, double* out, void* stream) {
  if (!v || !off || !out || nseg < 0) return MBPLS_ERR_ARG;
  if (nseg == 0) return MBPLS_OK;
  segsum_kernel<<<nseg, 1024, 0, static_cast<cudaStream_t>(stream)>>>(v, off, out);
  MBPLS_RETURN_LAST();
}

}  // extern "C"
