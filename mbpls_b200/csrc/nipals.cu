// NIPALS inner loop of multiblock PLS (mbpls/mbpls.py:809-993), dense and NaN-masked.
//
// One trip of the reference's `while diff_t > max_tol` loop (:841-914) is three launches:
//   xtu      : w~_j = x_j . u / u'u for every local feature j (+ per-CTA partial sums of w~_j^2 per block)
//   xw       : partial t~_b = sum_j w~_j x_j over a split of block b's features, for a chunk of samples
//   epilogue : fixed-order reduction of the partials, block-weight normalisation (deferred scalar),
//              superweights, superscore, convergence metric, Y weights and Y scores (:877-914)
// and one component is closed by `loadings_deflate` (:917-930, :968-969), which keeps each feature
// resident in shared memory between the loading dot product and the rank-1 update (1 read + 1 write),
// and can also emit the next component's first w~ (u restarts from the same Y column every
// component, :838, and Y is never deflated, :971-972).
//
// Every kernel starts with `if (*done) return;` so the host may enqueue trips in batches without
// reading the convergence flag after each one: trips launched after convergence are no-ops and the
// state left in device memory is exactly the state at loop exit.
#include "launch.cuh"
#include "../../include/mbpls_b200.h"

using namespace mbpls;

// Phase timestamps of xchg_epilogue_kernel for scripts/probes/xchg_probe.cu (compiled in with -DMBPLS_XCHG_STAMPS only)
#ifdef MBPLS_XCHG_STAMPS
__device__ unsigned long long g_xchg_stamps[16];
#define XSTAMP(k)                                                                   \
  do {                                                                              \
    if (threadIdx.x == 0) {                                                         \
      unsigned long long t_;                                                        \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                        \
      atomicMax(&g_xchg_stamps[k], t_);                                             \
    }                                                                               \
  } while (0)
#else
#define XSTAMP(k) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------
// xtu: one warp handles two features at a time, lanes stride along the sample axis with 16-byte
// loads, four loads per feature in flight.  NaN mode accumulates the masked denominator in the same
// pass (mbpls/mbpls.py:848-852); a column that turns out to be fully observed uses the plain u'u
// exactly like the reference's dense branch (:847).
// ------------------------------------------------------------------------------------------
template <bool NANMODE>
__device__ __forceinline__ void dot_accum(const double2 x, const double2 u, double& num, double& den, int& sawnan) {
  if (NANMODE) {
    const bool ox = !isnan(x.x), oy = !isnan(x.y);
    if (ox) { num = fma(x.x, u.x, num); den = fma(u.x, u.x, den); }
    if (oy) { num = fma(x.y, u.y, num); den = fma(u.y, u.y, den); }
    sawnan |= (!ox) | (!oy);
  } else {
    num = fma(x.x, u.x, num);
    num = fma(x.y, u.y, num);
  }
}

template <bool NANMODE>
__global__ void __launch_bounds__(256, NANMODE ? 2 : 0)
xtu_kernel(const double* __restrict__ Xt, long ld, int n, int p, int feats_per_cta, const double* __restrict__ u,
           const double* __restrict__ uu_ptr,  // plain denominator u'u; nullptr -> 1 (loadings, :920)
           const int* __restrict__ block_off, int B, double* __restrict__ w, double* __restrict__ norm_part,
           const int* __restrict__ done) {
  if (done && *done) return;
  extern __shared__ double wsm[];  // feats_per_cta doubles
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int f_begin = blockIdx.x * feats_per_cta;
  const int f_end = min(p, f_begin + feats_per_cta);
  const double uu = uu_ptr ? *uu_ptr : 1.0;
  const int n2 = n >> 1;
  const double2* __restrict__ u2 = reinterpret_cast<const double2*>(u);

  for (int j = f_begin + 2 * warp; j < f_end; j += 2 * nw) {
    const bool two = (j + 1 < f_end);
    const double2* __restrict__ x0 = reinterpret_cast<const double2*>(Xt + static_cast<size_t>(j) * ld);
    const double2* __restrict__ x1 = reinterpret_cast<const double2*>(Xt + static_cast<size_t>(two ? j + 1 : j) * ld);
    double na[2] = {0.0, 0.0}, nb[2] = {0.0, 0.0}, da[2] = {0.0, 0.0}, db[2] = {0.0, 0.0};
    int nan0 = 0, nan1 = 0;
    int i = lane;
    for (; i + 96 < n2; i += 128) {
      double2 a[4], b[4], uv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) a[k] = ld_stream(x0 + i + 32 * k);
#pragma unroll
      for (int k = 0; k < 4; ++k) b[k] = ld_stream(x1 + i + 32 * k);
#pragma unroll
      for (int k = 0; k < 4; ++k) uv[k] = u2[i + 32 * k];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        dot_accum<NANMODE>(a[k], uv[k], na[k & 1], da[k & 1], nan0);
        dot_accum<NANMODE>(b[k], uv[k], nb[k & 1], db[k & 1], nan1);
      }
    }
    for (; i < n2; i += 32) {
      const double2 uv = u2[i];
      dot_accum<NANMODE>(ld_stream(x0 + i), uv, na[0], da[0], nan0);
      dot_accum<NANMODE>(ld_stream(x1 + i), uv, nb[0], db[0], nan1);
    }
    if ((n & 1) && lane == 0) {  // odd tail element
      const double ut = u[n - 1];
      const double xa = Xt[static_cast<size_t>(j) * ld + n - 1];
      const double xb = Xt[static_cast<size_t>(two ? j + 1 : j) * ld + n - 1];
      dot_accum<NANMODE>(make_double2(xa, 0.0), make_double2(ut, 0.0), na[0], da[0], nan0);
      dot_accum<NANMODE>(make_double2(xb, 0.0), make_double2(ut, 0.0), nb[0], db[0], nan1);
    }
    const double num0 = warp_sum(na[0] + na[1]);
    const double num1 = warp_sum(nb[0] + nb[1]);
    double w0, w1;
    if (NANMODE) {
      const double den0 = warp_sum(da[0] + da[1]);
      const double den1 = warp_sum(db[0] + db[1]);
      nan0 = warp_or(nan0);
      nan1 = warp_or(nan1);
      w0 = nan0 ? num0 / den0 : num0 / uu;
      w1 = nan1 ? num1 / den1 : num1 / uu;
      if (!uu_ptr) {  // loadings: dense columns are not divided (:920), masked ones are (:923-925)
        w0 = nan0 ? num0 / den0 : num0;
        w1 = nan1 ? num1 / den1 : num1;
      }
    } else {
      w0 = uu_ptr ? num0 / uu : num0;
      w1 = uu_ptr ? num1 / uu : num1;
    }
    if (lane == 0) {
      w[j] = w0;
      wsm[j - f_begin] = w0;
      if (two) {
        w[j + 1] = w1;
        wsm[j + 1 - f_begin] = w1;
      }
    }
  }
  if (!norm_part) return;
  __syncthreads();
  // per-CTA partial of ||w~_b||^2 for every block, fixed order inside the CTA
  for (int b = warp; b < B; b += nw) {
    const int lo = max(block_off[b], f_begin), hi = min(block_off[b + 1], f_end);
    double s = 0.0;
    for (int j = lo + lane; j < hi; j += 32) {
      const double v = wsm[j - f_begin];
      s = fma(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) norm_part[static_cast<size_t>(blockIdx.x) * B + b] = s;
  }
}

// ------------------------------------------------------------------------------------------
// xw: each thread owns two consecutive samples; the CTA walks the features of its split eight at a
// time (eight independent 16-byte loads in flight per thread) and accumulates w~_j * x_ij in
// registers.  NaN mode also accumulates the masked sum of w~_j^2 (mbpls/mbpls.py:867-872).
// grid = (sample chunks of 512, splits); partials go to Tnum[split][ld] (and Tden).
// ------------------------------------------------------------------------------------------
template <bool NANMODE>
__global__ void __launch_bounds__(256, 4)
xw_kernel(const double* __restrict__ Xt, long ld, int n, const double* __restrict__ w,
          const int* __restrict__ split_f0, const int* __restrict__ split_f1, double* __restrict__ Tnum,
          double* __restrict__ Tden, long ldt, const int* __restrict__ done) {
  if (done && *done) return;
  const int s = blockIdx.y;
  const int f0 = split_f0[s], f1 = split_f1[s];
  const int r = (blockIdx.x * 256 + threadIdx.x) * 2;
  if (r >= n) return;
  const double* __restrict__ xp = Xt + r;
  double nx = 0.0, ny = 0.0, dx = 0.0, dy = 0.0;
  int j = f0;
  for (; j + 8 <= f1; j += 8) {
    double2 x[8];
    double wj[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = ld_stream(reinterpret_cast<const double2*>(xp + static_cast<size_t>(j + k) * ld));
#pragma unroll
    for (int k = 0; k < 8; ++k) wj[k] = __ldg(w + j + k);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (NANMODE) {
        const bool ox = !isnan(x[k].x), oy = !isnan(x[k].y);
        const double w2 = wj[k] * wj[k];
        nx = fma(ox ? x[k].x : 0.0, wj[k], nx);
        ny = fma(oy ? x[k].y : 0.0, wj[k], ny);
        dx += ox ? w2 : 0.0;
        dy += oy ? w2 : 0.0;
      } else {
        nx = fma(x[k].x, wj[k], nx);
        ny = fma(x[k].y, wj[k], ny);
      }
    }
  }
  for (; j < f1; ++j) {
    const double2 x = ld_stream(reinterpret_cast<const double2*>(xp + static_cast<size_t>(j) * ld));
    const double wj = __ldg(w + j);
    if (NANMODE) {
      const bool ox = !isnan(x.x), oy = !isnan(x.y);
      const double w2 = wj * wj;
      nx = fma(ox ? x.x : 0.0, wj, nx);
      ny = fma(oy ? x.y : 0.0, wj, ny);
      dx += ox ? w2 : 0.0;
      dy += oy ? w2 : 0.0;
    } else {
      nx = fma(x.x, wj, nx);
      ny = fma(x.y, wj, ny);
    }
  }
  double* tn = Tnum + static_cast<size_t>(s) * ldt + r;
  *reinterpret_cast<double2*>(tn) = make_double2(nx, ny);  // r is even and r+1 < ldt (ld % 16 == 0)
  if (NANMODE) {
    double* td = Tden + static_cast<size_t>(s) * ldt + r;
    *reinterpret_cast<double2*>(td) = make_double2(dx, dy);
  }
}

// ------------------------------------------------------------------------------------------
// reduce_partials: red[b][i] = sum over the splits of block b (ascending split index) of Tnum, same
// for Tden; red tail holds ||w~_b||^2 = sum over xtu CTAs (ascending) of norm_part.  `red` is the
// buffer that is all-reduced across GPUs when features are sharded.
// layout of red: [B][ldt] numerators | [B][ldt] denominators (NaN mode only) | [B] squared norms
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const double* __restrict__ Tnum, const double* __restrict__ Tden, long ldt, int n, int B,
                       const int* __restrict__ block_split_off, const double* __restrict__ norm_part, int n_norm_parts,
                       double* __restrict__ red, int nanmode, const int* __restrict__ done) {
  if (done && *done) return;
  const int b = blockIdx.y;
  const int s0 = block_split_off[b], s1 = block_split_off[b + 1];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double a = 0.0, d = 0.0;
    for (int s = s0; s < s1; ++s) a += Tnum[static_cast<size_t>(s) * ldt + i];
    red[static_cast<size_t>(b) * ldt + i] = a;
    if (nanmode) {
      for (int s = s0; s < s1; ++s) d += Tden[static_cast<size_t>(s) * ldt + i];
      red[static_cast<size_t>(B + b) * ldt + i] = d;
    }
  }
  if (blockIdx.x == 0) {  // squared block-weight norm: fixed-order tree over the xtu CTAs
    __shared__ double scratch[32];
    double acc = 0.0;
    for (int c = threadIdx.x; c < n_norm_parts; c += blockDim.x) acc += norm_part[static_cast<size_t>(c) * B + b];
    acc = block_sum1(acc, scratch);
    if (threadIdx.x == 0) red[static_cast<size_t>(nanmode ? 2 * B : B) * ldt + b] = acc;
  }
}

// ------------------------------------------------------------------------------------------
// begin_component: u <- u0 (mbpls/mbpls.py:832-838), u'u, trip counter 0, diff 1, done 0.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
begin_component_kernel(const double* __restrict__ u0, int n, double* __restrict__ u, double* __restrict__ scal,
                       int* __restrict__ ctrl) {
  __shared__ double scratch[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = u0[i];
    u[i] = v;
    acc = fma(v, v, acc);
  }
  acc = block_sum1(acc, scratch);
  if (threadIdx.x == 0) {
    scal[MBPLS_SCAL_UU] = acc;
    scal[MBPLS_SCAL_DIFF] = 1.0;
    ctrl[MBPLS_CTRL_DONE] = 0;
    ctrl[MBPLS_CTRL_TRIPS] = 0;
  }
}

// ------------------------------------------------------------------------------------------
// epilogue: single CTA, fixed-order reductions -> bitwise reproducible on every GPU.
// ------------------------------------------------------------------------------------------
#define EPI_MAXB 64
#define EPI_MAXQ 64

__device__ __forceinline__ void epilogue_body(const mbpls_epilogue_args& a) {
  int* ctrl = a.ctrl;
  __shared__ double scratch[32 * 4];
  __shared__ double s_norm[EPI_MAXB], s_a[EPI_MAXB], s_v[EPI_MAXQ];
  __shared__ double s_sc[8];
  const int n = a.n, B = a.B, q = a.q;
  const long ldt = a.ldt;
  const int tid = threadIdx.x, nt = blockDim.x;
  const int nan = a.nanmode;
  const double* red_num = a.red;
  const double* red_den = a.red + static_cast<size_t>(B) * ldt;
  const double* red_nrm = a.red + static_cast<size_t>(nan ? 2 * B : B) * ldt;
  const double uu = a.scal[MBPLS_SCAL_UU];

  if (tid < B) s_norm[tid] = sqrt(red_nrm[tid]);
  __syncthreads();

  // phase 1: block scores t_b (:862-875) and T'u (:879)
  for (int b = 0; b < B; ++b) {
    const double nb = s_norm[b];
    const UniformDivisor by_nb(nb);  // (bit-identical to `/ nb`, a tenth of the instructions: common.cuh)
    double acc = 0.0;
    double* tb = a.T + static_cast<size_t>(b) * ldt;
    for (int i = tid; i < n; i += nt) {
      const double num = red_num[static_cast<size_t>(b) * ldt + i];
      double t;
      if (nan && a.row_flag[static_cast<size_t>(b) * a.ldf + i]) t = nb * num / red_den[static_cast<size_t>(b) * ldt + i];
      else t = by_nb(num);
      tb[i] = t;
      acc = fma(t, a.u[i], acc);
    }
    acc = block_sum1(acc, scratch);
    if (tid == 0) s_a[b] = acc / uu;
  }
  __syncthreads();
  if (tid == 0) {  // superweights to unit length (:880)
    double s = 0.0;
    for (int b = 0; b < B; ++b) s = fma(s_a[b], s_a[b], s);
    s = sqrt(s);
    for (int b = 0; b < B; ++b) {
      s_a[b] /= s;
      a.a[b] = s_a[b];
    }
  }
  __syncthreads();

  // phase 2: superscore ts = T a, normalised (:882-883)
  double ss = 0.0;
  for (int i = tid; i < n; i += nt) {
    double t = 0.0;
    for (int b = 0; b < B; ++b) t = fma(a.T[static_cast<size_t>(b) * ldt + i], s_a[b], t);
    a.ts[i] = t;
    ss = fma(t, t, ss);
  }
  ss = block_sum1(ss, scratch);
  const double tsn = sqrt(ss);
  const UniformDivisor by_tsn(tsn);

  // phase 3: normalise, convergence metric against ts_old (:884-888), ts'ts
  double v4[4] = {0.0, 0.0, 0.0, 0.0};  // sum d^2, sum |d|, (unused), ts'ts
  double dmax = 0.0, dmin = INFINITY;
  for (int i = tid; i < n; i += nt) {
    const double t = by_tsn(a.ts[i]);
    const double d = a.ts_old[i] - t;
    a.ts[i] = t;
    a.ts_old[i] = t;
    v4[0] = fma(d, d, v4[0]);
    v4[1] += fabs(d);
    dmax = fmax(dmax, fabs(d));
    dmin = fmin(dmin, fabs(d));
    v4[3] = fma(t, t, v4[3]);
  }
  block_sum<4>(v4, scratch);
  dmax = warp_max(dmax);
  dmin = warp_min(dmin);
  __syncthreads();
  if ((tid & 31) == 0) {
    scratch[tid >> 5] = dmax;
    scratch[32 + (tid >> 5)] = dmin;
  }
  __syncthreads();
  if (tid == 0) {
    double mx = 0.0, mn = INFINITY;
    for (int wv = 0; wv < (nt >> 5); ++wv) {
      mx = fmax(mx, scratch[wv]);
      mn = fmin(mn, scratch[32 + wv]);
    }
    double diff;
    switch (a.norm_kind) {
      case MBPLS_NORM_L1: diff = v4[1]; break;
      case MBPLS_NORM_MAX: diff = mx; break;
      case MBPLS_NORM_MIN: diff = mn; break;
      default: diff = sqrt(v4[0]);
    }
    s_sc[0] = diff;
  }
  __syncthreads();
  const double tt = v4[3];

  // phase 4: Y weights v = Y'ts / ts'ts (:890-899), masked for Y columns with NaN
  for (int c = 0; c < q; ++c) {
    const double* y = a.Yt + static_cast<size_t>(c) * ldt;
    double nd[2] = {0.0, 0.0};
    const bool masked = nan && a.ycol_flag[c];
    for (int i = tid; i < n; i += nt) {
      const double yi = y[i], t = a.ts[i];
      if (masked) {
        if (!isnan(yi)) {
          nd[0] = fma(yi, t, nd[0]);
          nd[1] = fma(t, t, nd[1]);
        }
      } else {
        nd[0] = fma(yi, t, nd[0]);
      }
    }
    block_sum<2>(nd, scratch);
    if (tid == 0) s_v[c] = masked ? nd[0] / nd[1] : nd[0] / tt;
  }
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int c = 0; c < q; ++c) {
      s = fma(s_v[c], s_v[c], s);
      a.v[c] = s_v[c];
    }
    s_sc[1] = s;
  }
  __syncthreads();
  const double vv = s_sc[1];

  // phase 5: Y scores u = Y v / v'v, normalised (:901-913).  NaN mode: rows flagged in the *last X
  // block* use observed Y columns only (the reference indexes sparse_X_info_[block] at :903).
  double un = 0.0;
  const UniformDivisor by_vv(vv);
  for (int i = tid; i < n; i += nt) {
    double val;
    if (nan && a.row_flag[static_cast<size_t>(B - 1) * a.ldf + i]) {
      double num = 0.0, den = 0.0;
      for (int c = 0; c < q; ++c) {
        const double yi = a.Yt[static_cast<size_t>(c) * ldt + i];
        if (!isnan(yi)) {
          num = fma(yi, s_v[c], num);
          den = fma(s_v[c], s_v[c], den);
        }
      }
      val = num / den;
    } else {
      double num = 0.0;
      for (int c = 0; c < q; ++c) num = fma(a.Yt[static_cast<size_t>(c) * ldt + i], s_v[c], num);
      val = by_vv(num);
    }
    a.u[i] = val;
    un = fma(val, val, un);
  }
  un = block_sum1(un, scratch);
  const double unorm = sqrt(un);
  const UniformDivisor by_un(unorm);
  double uu_new = 0.0;
  for (int i = tid; i < n; i += nt) {
    const double val = by_un(a.u[i]);
    a.u[i] = val;
    uu_new = fma(val, val, uu_new);
  }
  uu_new = block_sum1(uu_new, scratch);
  if (tid == 0) {
    const int trips = ctrl[MBPLS_CTRL_TRIPS] + 1;
    ctrl[MBPLS_CTRL_TRIPS] = trips;
    a.scal[MBPLS_SCAL_UU] = uu_new;
    a.scal[MBPLS_SCAL_TT] = tt;
    a.scal[MBPLS_SCAL_VV] = vv;
    if (trips > 1) {  // the first trip has nothing to compare with (:884-885)
      a.scal[MBPLS_SCAL_DIFF] = s_sc[0];
      if (!(s_sc[0] > a.max_tol)) ctrl[MBPLS_CTRL_DONE] = 1;
    }
    if (a.diff_trace && trips <= a.diff_trace_len) a.diff_trace[trips - 1] = trips > 1 ? s_sc[0] : 1.0;
  }
}

// The same step for the common case -- dense data, at most EPS_ITEMS * 1024 samples, at most 8 blocks, at most 16 Y columns --
// with every thread keeping its samples of u and ts in registers: the B block passes become ONE pass with 8 accumulators and
// one 8-wide block reduction, the q Y columns likewise, and no vector is re-read from global memory between the phases.
// (The general body above makes 9 latency-bound passes and 9 reductions for B = 4, q = 1: 59 us at n = 10,000, which is a
// third of a PLS2 trip on an 8-GPU shard.)  Same arithmetic per element; only the order of the block-wide sums differs.
// Every phase first issues ALL of its global loads into registers and only then computes and stores: the stores (T, ts_old)
// may alias the loads as far as the compiler can tell, and interleaved they ran as ~40 dependent L2 round trips (20 us for the
// block scores alone, scripts/probes/xchg_probe.cu).
#define EPS_ITEMS 10
__device__ __forceinline__ bool epilogue_small_ok(const mbpls_epilogue_args& a) {
  return !a.nanmode && a.n <= EPS_ITEMS * 1024 && a.B <= 8 && a.q <= 16 && blockDim.x == 1024;
}

__device__ __forceinline__ void epilogue_body_small(const mbpls_epilogue_args& a) {
  int* ctrl = a.ctrl;
  __shared__ double scratch[32 * 8];
  __shared__ double s_norm[8], s_a[8], s_v[16];
  __shared__ double s_sc[8];
  const int n = a.n, B = a.B, q = a.q;
  const long ldt = a.ldt;
  const int tid = threadIdx.x;
  constexpr int NT = 1024;
  const double* red_num = a.red;
  const double* red_nrm = a.red + static_cast<size_t>(B) * ldt;
  const double uu = a.scal[MBPLS_SCAL_UU];
  if (tid < B) s_norm[tid] = sqrt(red_nrm[tid]);
  double ur[EPS_ITEMS];
#pragma unroll
  for (int k = 0; k < EPS_ITEMS; ++k) {
    const int i = tid + k * NT;
    ur[k] = i < n ? a.u[i] : 0.0;
  }
  __syncthreads();

  // phase 1: block scores t_b (:862-875) and T'u (:879), all blocks in one pass
  double ab[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) ab[b] = 0.0;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    if (b < B) {
      const UniformDivisor by_nb(s_norm[b]);  // (bit-identical to `/ nb`, a tenth of the instructions: common.cuh)
      double num[EPS_ITEMS];
#pragma unroll
      for (int k = 0; k < EPS_ITEMS; ++k) {
        const int i = tid + k * NT;
        num[k] = i < n ? __ldcg(red_num + static_cast<size_t>(b) * ldt + i) : 0.0;
      }
#pragma unroll
      for (int k = 0; k < EPS_ITEMS; ++k) {
        const int i = tid + k * NT;
        if (i < n) {
          const double t = by_nb(num[k]);
          a.T[static_cast<size_t>(b) * ldt + i] = t;
          ab[b] = fma(t, ur[k], ab[b]);
        }
      }
    }
  }
  block_sum<8>(ab, scratch);
  XSTAMP(5);
  if (tid == 0) {  // superweights to unit length (:880)
    double s = 0.0;
    for (int b = 0; b < B; ++b) {
      s_a[b] = ab[b] / uu;
      s = fma(s_a[b], s_a[b], s);
    }
    s = sqrt(s);
    for (int b = 0; b < B; ++b) {
      s_a[b] /= s;
      a.a[b] = s_a[b];
    }
  }
  __syncthreads();

  // phase 2: superscore ts = T a, normalised (:882-883)
  double tsr[EPS_ITEMS], old[EPS_ITEMS];
  double ss = 0.0;
#pragma unroll
  for (int k = 0; k < EPS_ITEMS; ++k) {
    const int i = tid + k * NT;
    tsr[k] = 0.0;
    old[k] = i < n ? a.ts_old[i] : 0.0;  // for phase 3: in flight during the reduction below
  }
  for (int b = 0; b < B; ++b) {  // same order of the sum over blocks per element as before: b ascending
    const double ab_ = s_a[b];
#pragma unroll
    for (int k = 0; k < EPS_ITEMS; ++k) {
      const int i = tid + k * NT;
      if (i < n) tsr[k] = fma(a.T[static_cast<size_t>(b) * ldt + i], ab_, tsr[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < EPS_ITEMS; ++k) ss = fma(tsr[k], tsr[k], ss);
  ss = block_sum1(ss, scratch);
  const double tsn = sqrt(ss);
  const UniformDivisor by_tsn(tsn);

  // phase 3: normalise, convergence metric against ts_old (:884-888), ts'ts
  double v4[4] = {0.0, 0.0, 0.0, 0.0};
  double dmax = 0.0, dmin = INFINITY;
#pragma unroll
  for (int k = 0; k < EPS_ITEMS; ++k) {
    const int i = tid + k * NT;
    if (i < n) {
      const double t = by_tsn(tsr[k]);
      const double d = old[k] - t;
      tsr[k] = t;
      a.ts[i] = t;
      a.ts_old[i] = t;
      v4[0] = fma(d, d, v4[0]);
      v4[1] += fabs(d);
      dmax = fmax(dmax, fabs(d));
      dmin = fmin(dmin, fabs(d));
      v4[3] = fma(t, t, v4[3]);
    }
  }
  block_sum<4>(v4, scratch);
  dmax = warp_max(dmax);
  dmin = warp_min(dmin);
  __syncthreads();
  if ((tid & 31) == 0) {
    scratch[tid >> 5] = dmax;
    scratch[32 + (tid >> 5)] = dmin;
  }
  __syncthreads();
  if (tid == 0) {
    double mx = 0.0, mn = INFINITY;
    for (int wv = 0; wv < 32; ++wv) {
      mx = fmax(mx, scratch[wv]);
      mn = fmin(mn, scratch[32 + wv]);
    }
    double diff;
    switch (a.norm_kind) {
      case MBPLS_NORM_L1: diff = v4[1]; break;
      case MBPLS_NORM_MAX: diff = mx; break;
      case MBPLS_NORM_MIN: diff = mn; break;
      default: diff = sqrt(v4[0]);
    }
    s_sc[0] = diff;
  }
  const double tt = v4[3];
  XSTAMP(6);

  // phase 4: Y weights v = Y'ts / ts'ts (:899), eight columns per pass
  for (int c0 = 0; c0 < q; c0 += 8) {
    double vb[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) vb[c] = 0.0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      if (c0 + c < q) {
        const double* y = a.Yt + static_cast<size_t>(c0 + c) * ldt;
#pragma unroll
        for (int k = 0; k < EPS_ITEMS; ++k) {
          const int i = tid + k * NT;
          if (i < n) vb[c] = fma(y[i], tsr[k], vb[c]);
        }
      }
    }
    block_sum<8>(vb, scratch);
    if (tid == 0)
      for (int c = 0; c < 8 && c0 + c < q; ++c) s_v[c0 + c] = vb[c] / tt;
  }
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    for (int c = 0; c < q; ++c) {
      s = fma(s_v[c], s_v[c], s);
      a.v[c] = s_v[c];
    }
    s_sc[1] = s;
  }
  __syncthreads();
  const double vv = s_sc[1];

  // phase 5: Y scores u = Y v / v'v, normalised (:911-913)
  XSTAMP(7);
  double un = 0.0;
  const UniformDivisor by_vv(vv);
#pragma unroll
  for (int k = 0; k < EPS_ITEMS; ++k) ur[k] = 0.0;
  for (int c = 0; c < q; ++c) {  // (column loop outside: the loads of one column are independent of each other)
    const double vc = s_v[c];
    const double* y = a.Yt + static_cast<size_t>(c) * ldt;
#pragma unroll
    for (int k = 0; k < EPS_ITEMS; ++k) {
      const int i = tid + k * NT;
      if (i < n) ur[k] = fma(y[i], vc, ur[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < EPS_ITEMS; ++k) {
    const int i = tid + k * NT;
    const double val = i < n ? by_vv(ur[k]) : 0.0;
    ur[k] = val;
    un = fma(val, val, un);
  }
  un = block_sum1(un, scratch);
  const double unorm = sqrt(un);
  const UniformDivisor by_un(unorm);
  double uu_new = 0.0;
#pragma unroll
  for (int k = 0; k < EPS_ITEMS; ++k) {
    const int i = tid + k * NT;
    if (i < n) {
      const double val = by_un(ur[k]);
      a.u[i] = val;
      uu_new = fma(val, val, uu_new);
    }
  }
  uu_new = block_sum1(uu_new, scratch);
  if (tid == 0) {
    const int trips = ctrl[MBPLS_CTRL_TRIPS] + 1;
    ctrl[MBPLS_CTRL_TRIPS] = trips;
    a.scal[MBPLS_SCAL_UU] = uu_new;
    a.scal[MBPLS_SCAL_TT] = tt;
    a.scal[MBPLS_SCAL_VV] = vv;
    if (trips > 1) {  // the first trip has nothing to compare with (:884-885)
      a.scal[MBPLS_SCAL_DIFF] = s_sc[0];
      if (!(s_sc[0] > a.max_tol)) ctrl[MBPLS_CTRL_DONE] = 1;
    }
    if (a.diff_trace && trips <= a.diff_trace_len) a.diff_trace[trips - 1] = trips > 1 ? s_sc[0] : 1.0;
  }
}

__global__ void __launch_bounds__(1024) nipals_epilogue_kernel(mbpls_epilogue_args a) {
  if (a.ctrl[MBPLS_CTRL_DONE]) return;
  if (epilogue_small_ok(a)) epilogue_body_small(a);
  else epilogue_body(a);
}

// ------------------------------------------------------------------------------------------
// xchg_epilogue: reduce_partials + the exchange between the GPUs + the epilogue as ONE kernel.
//
// Feature-sharded fits end every trip with "sum the split partials -> sum over the GPUs -> superlevel step on every GPU".
// With NCCL that is reduce_partials, a copy, the all-reduce on NCCL's stream (two cross-stream event hand-offs) and the
// epilogue: four launches and ~0.1 ms of fixed cost per trip, which is what 8 GPUs lose on the headline problem and most of
// a PLS2 trip on a small shard.  Here the exchange runs over NVLink peer memory inside the kernel:
//   A  every CTA sums the split partials of its items (sample i of block b) into THIS rank's symmetric buffer (slot = trip
//      parity); the last CTA to finish publishes `seq` into the flag word "rank -> peer" of every peer's buffer
//      (st.release.sys over NVLink) and into its own;
//   B  every CTA waits until all `world` flags of its own buffer have reached `seq` (ld.acquire.sys), then adds the slots of
//      all ranks IN RANK ORDER -- peer slots are read straight through NVLink -- so every GPU forms bit-identical sums, which
//      keeps the replicated epilogues (and their convergence decisions) in lock step;
//   C  the last CTA to finish B runs the epilogue on the sums.
// Slots alternate with the trip parity: a rank can only start trip s+2 after every peer has published s+1, which it does
// after it has finished reading the slots of trip s.  A rank that waits longer than XCHG_TIMEOUT_NS for a peer raises
// ctrl[MBPLS_CTRL_ERROR] instead of hanging.  world == 1: A writes the sums directly, C follows -- one launch instead of two.
// ------------------------------------------------------------------------------------------
#define XCHG_TIMEOUT_NS 20000000000ull

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(1024) xchg_epilogue_kernel(mbpls_xchg_args x) {
  const mbpls_epilogue_args& a = x.epi;
  if (a.ctrl[MBPLS_CTRL_DONE]) return;
  __shared__ int s_flag;
  const int n = a.n, B = a.B, nan = a.nanmode;
  const long ldt = a.ldt;
  const int nrows = nan ? 2 * B : B;                    // rows of `red` that hold per-sample sums
  const long nitems = static_cast<long>(nrows) * n;
  const long norm_off = static_cast<long>(nrows) * ldt;  // B squared norms behind them
  const int tid = threadIdx.x, G = gridDim.x;
  const bool multi = x.world > 1;
  double* mine = multi ? reinterpret_cast<double*>(x.peer_bufs[x.rank]) + (x.seq & 1ull) * x.slot_elems : const_cast<double*>(a.red);
  XSTAMP(0);

  // ---- A: split partials -> this rank's sums.  An item (row r, sample i) is the sum of its block's split partials; `tpi`
  // warps share an item (warp w of a group takes the splits s0 + w, s0 + w + tpi, ...; lanes = 32 consecutive samples, so every
  // load is a full 256-byte row segment) and their partial sums are added in a fixed order -- short problems with many splits
  // (n = 2,000 in 592 splits: 222 dependent additions per item) would otherwise leave most of the CTAs idle.
  {
    __shared__ double s_part[1024];
    const int tpi = x.tpi > 1 ? x.tpi : 1;  // power of two <= 32, the same for every launch of a fit
    const int ipc = 1024 / tpi;             // items per CTA and round
    const int warp = tid >> 5, lane = tid & 31;
    const int sub = warp % tpi, item_local = (warp / tpi) * 32 + lane;
    for (long base = static_cast<long>(blockIdx.x) * ipc; base < nitems; base += static_cast<long>(G) * ipc) {
      const long it = base + item_local;
      double acc = 0.0;
      int r = 0, i = 0;
      if (it < nitems) {
        r = static_cast<int>(it / n);
        i = static_cast<int>(it - static_cast<long>(r) * n);
        const int b = r < B ? r : r - B;
        const double* src = r < B ? x.Tnum : x.Tden;
        const int s0 = x.block_split_off[b], s1 = x.block_split_off[b + 1];
        for (int sp = s0 + sub; sp < s1; sp += tpi) acc += src[static_cast<size_t>(sp) * x.ldp + i];
      }
      if (tpi == 1) {
        if (it < nitems) mine[static_cast<size_t>(r) * ldt + i] = acc;
      } else {
        s_part[sub * ipc + item_local] = acc;
        __syncthreads();
        if (tid < ipc && base + tid < nitems) {
          const long it2 = base + tid;
          const int r2 = static_cast<int>(it2 / n), i2 = static_cast<int>(it2 - static_cast<long>(r2) * n);
          double t = 0.0;
          for (int k = 0; k < tpi; ++k) t += s_part[k * ipc + tid];
          mine[static_cast<size_t>(r2) * ldt + i2] = t;
        }
        __syncthreads();
      }
    }
  }
  if (blockIdx.x == 0 && tid < 32) {  // squared block-weight norms: one lane per block at a time, fixed order
    for (int b = tid; b < B; b += 32) {
      double acc = 0.0;
      for (int c = 0; c < x.n_norm_parts; ++c) acc += x.norm_part[static_cast<size_t>(c) * B + b];
      mine[norm_off + b] = acc;
    }
  }
  XSTAMP(1);
  if (multi) {
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      const unsigned old = atomicAdd(&x.counters[0], 1u);
      if (old == static_cast<unsigned>(G) - 1u) {  // all CTAs of this rank have written their sums: publish
        x.counters[0] = 0u;
        __threadfence_system();
        for (int r = 0; r < x.world; ++r) {
          unsigned long long* flags = reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(x.peer_bufs[r]) + x.flags_off);
          st_release_sys_u64(flags + x.rank, x.seq);
        }
      }
      // ---- B: wait for every rank's sums (this rank's included)
      const unsigned long long* myflags =
          reinterpret_cast<const unsigned long long*>(reinterpret_cast<double*>(x.peer_bufs[x.rank]) + x.flags_off);
      const unsigned long long t0 = global_timer_ns();
      int ok = 1;
      for (int r = 0; r < x.world && ok; ++r) {
        while (ld_acquire_sys_u64(myflags + r) < x.seq) {
          if (global_timer_ns() - t0 > XCHG_TIMEOUT_NS) {
            ok = 0;
            break;
          }
        }
      }
      s_flag = ok;
    }
    __syncthreads();
    if (!s_flag) {  // a peer never arrived: report instead of hanging (the host raises)
      if (tid == 0) a.ctrl[MBPLS_CTRL_ERROR] = 1;
      return;
    }
    XSTAMP(2);
    double* red = const_cast<double*>(a.red);
    const size_t slot = (x.seq & 1ull) * x.slot_elems;
    for (long it = static_cast<long>(blockIdx.x) * blockDim.x + tid; it < nitems + B; it += static_cast<long>(G) * blockDim.x) {
      size_t off;
      if (it < nitems) {
        const int r = static_cast<int>(it / n), i = static_cast<int>(it - static_cast<long>(r) * n);
        off = static_cast<size_t>(r) * ldt + i;
      } else {
        off = static_cast<size_t>(norm_off) + (it - nitems);
      }
      double acc = 0.0;
      for (int r = 0; r < x.world; ++r) acc += __ldcv(reinterpret_cast<const double*>(x.peer_bufs[r]) + slot + off);
      red[off] = acc;
    }
  }
  // ---- C: the last CTA runs the superlevel step on the sums
  XSTAMP(3);
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned old = atomicAdd(&x.counters[1], 1u);
    s_flag = old == static_cast<unsigned>(G) - 1u;
    if (s_flag) x.counters[1] = 0u;
  }
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  XSTAMP(4);
  if (epilogue_small_ok(a)) epilogue_body_small(a);
  else epilogue_body(a);
  XSTAMP(9);
}

// ------------------------------------------------------------------------------------------
// xchg_epilogue_mc: the same step with the superlevel arithmetic spread over ALL CTAs of the launch.
//
// One CTA doing the superlevel step moves ~1.5 MB through a single SM and makes ~10 dependent trips to L2 / HBM (everything small
// has been evicted by the pass over X that precedes it): 50-65 us inside a fit (scripts/xchg_stamps.py).  Here CTA c owns the
// samples [c * ch, (c + 1) * ch): it forms their split sums,
// keeps block scores, superscore and Y rows of its samples in registers (one sample per thread), and
// the five sums over all samples -- T'u, |ts|, {diff_t, ts'ts, Y'ts}, |u|, u'u -- are grid-wide reductions: every CTA writes its
// partial into `work` and releases a flag, every CTA acquires all G flags (all CTAs are resident) and adds the G partials in
// the same fixed order, so all of them -- and all GPUs -- continue from bit-identical scalars.
// Each grid-wide sum costs ~5 us (release, acquire and read are three trips to L2), so the step takes 50 instead of 65 us
// (PLS2 trip of 8 blocks, n = 2,000: 85 instead of 99 us for the whole launch).  Single GPU, dense data, B <= 8, q <= 16,
// ch <= 1024 (n <= 1024 * CTAs); everything else takes xchg_epilogue_kernel -- between GPUs the launch is dominated by the wait
// for the slowest rank's trip kernel and the two forms measured the same (2 GPUs: 109.4 vs 108.6 ms per fit).
// ------------------------------------------------------------------------------------------
#define XMC_W 32       // doubles per (phase, CTA) slot of the work buffer
#define XMC_GMAX 160   // slots per phase

__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Grid-wide sum of NV values per CTA (all CTAs of the launch are resident).  Thread 0 writes this CTA's partials into its slot
// of `work` and then RELEASES the launch's epoch into its flag word; thread g < G ACQUIRES CTA g's flag (one poll per CTA, all
// in flight at once: no atomics, no arrival counter, no generation word -- the first version, an atomic-counter barrier, cost
// ~6 us per sum, four dependent L2 round trips); after the CTA barrier warp k adds value k over the CTAs -- lanes take CTAs
// lane, lane + 32, ... in order, then the fixed shuffle tree -- so every CTA ends up with identical bits in s_out[0 .. NV).
// Flags only ever grow (epoch = number of the launch within the fit), one flag array per phase.
template <int NV>
__device__ __forceinline__ void grid_sum(double (&v)[NV], double* scratch, double* s_out, double* work, int phase,
                                         const mbpls_xchg_args& x, int G) {
  block_sum<NV>(v, scratch);
  unsigned long long* flags = reinterpret_cast<unsigned long long*>(work + static_cast<size_t>(6) * XMC_GMAX * XMC_W) + phase * XMC_GMAX;
  if (threadIdx.x == 0) {
    double* slot = work + (static_cast<size_t>(phase) * XMC_GMAX + blockIdx.x) * XMC_W;
#pragma unroll
    for (int k = 0; k < NV; ++k) slot[k] = v[k];
    st_release_gpu_u64(flags + blockIdx.x, x.epoch);
  }
  if (threadIdx.x < G) {
    while (ld_acquire_gpu_u64(flags + threadIdx.x) < x.epoch) {
    }
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp < NV) {
    const double* base = work + static_cast<size_t>(phase) * XMC_GMAX * XMC_W + warp;
    double t = 0.0;
    for (int g = lane; g < G; g += 32) t += __ldcg(base + static_cast<size_t>(g) * XMC_W);
    t = warp_sum(t);
    if (lane == 0) s_out[warp] = t;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024) xchg_epilogue_mc_kernel(mbpls_xchg_args x) {
  const mbpls_epilogue_args& a = x.epi;
  if (a.ctrl[MBPLS_CTRL_DONE]) return;
  __shared__ double s_part[1024];
  __shared__ double scratch[32 * 19];
  __shared__ double s_tot[32];
  __shared__ double s_norm[8], s_a[8], s_v[16];
  __shared__ double s_ext[64];
  const int n = a.n, B = a.B, q = a.q;
  const long ldt = a.ldt;
  const int tid = threadIdx.x, G = gridDim.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int ch = x.ch;                              // samples per CTA, a multiple of 32
  const int i0 = blockIdx.x * ch;
  const long norm_off = static_cast<long>(B) * ldt;  // B squared norms behind the B x ldt sums
  double* red = const_cast<double*>(a.red);
  double* mine = red;  // (single GPU: the sums go straight into red)
  XSTAMP(0);

  // ---- A: split partials of this CTA's samples -> sums (tpi warps per item, see xchg_epilogue_kernel)
  {
    const int tpi = x.tpi > 1 ? x.tpi : 1;
    const int ipc = 1024 / tpi;
    const int sub = warp % tpi, l0 = (warp / tpi) * 32 + lane;
    const int nloc = B * ch;  // local items: block b, local sample il
    for (int base = 0; base < nloc; base += ipc) {
      const int l = base + l0;
      const int b = l / ch, il = l - b * ch, i = i0 + il;
      double acc = 0.0;
      const bool live = l < nloc && i < n;
      if (live) {
        const int s0 = x.block_split_off[b], s1 = x.block_split_off[b + 1];
        for (int sp = s0 + sub; sp < s1; sp += tpi) acc += x.Tnum[static_cast<size_t>(sp) * x.ldp + i];
      }
      if (tpi == 1) {
        if (live) mine[static_cast<size_t>(b) * ldt + i] = acc;
      } else {
        s_part[sub * ipc + l0] = acc;
        __syncthreads();
        if (tid < ipc) {
          const int l2 = base + tid;
          const int b2 = l2 / ch, i2 = i0 + (l2 - b2 * ch);
          if (l2 < nloc && i2 < n) {
            double t = 0.0;
            for (int k = 0; k < tpi; ++k) t += s_part[k * ipc + tid];
            mine[static_cast<size_t>(b2) * ldt + i2] = t;
          }
        }
        __syncthreads();
      }
    }
  }
  // squared block-weight norms: warp b adds the parts lane, lane + 32, ... in order, then the shuffle tree; every CTA forms
  // them (identical bits everywhere)
  if (warp < B) {
    double t = 0.0;
    for (int c = lane; c < x.n_norm_parts; c += 32) t += x.norm_part[static_cast<size_t>(c) * B + warp];
    t = warp_sum(t);
    if (lane == 0) s_ext[warp] = t;
  }
  XSTAMP(1);
  __syncthreads();
  if (blockIdx.x == 0 && tid < B) red[norm_off + tid] = s_ext[tid];  // record_component reads the norms from red
  if (tid < B) s_norm[tid] = sqrt(s_ext[tid]);
  __syncthreads();
  XSTAMP(3);
  XSTAMP(4);

  // ---- C: superlevel step; thread tid < ch owns sample i0 + tid
  const int i = i0 + tid;
  const bool own = tid < ch && i < n;
  const double uu = a.scal[MBPLS_SCAL_UU];
  double tb[8], yv[16];
  double ui = 0.0, old = 0.0;
#pragma unroll
  for (int b = 0; b < 8; ++b) tb[b] = (own && b < B) ? __ldcg(red + static_cast<size_t>(b) * ldt + i) : 0.0;
#pragma unroll
  for (int c = 0; c < 16; ++c) yv[c] = (own && c < q) ? a.Yt[static_cast<size_t>(c) * ldt + i] : 0.0;
  if (own) {
    ui = a.u[i];
    old = a.ts_old[i];
  }
  // phase 1: block scores t_b (:862-875) and T'u (:879)
  double ab[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    ab[b] = 0.0;
    if (b < B) {
      const UniformDivisor by_nb(s_norm[b]);
      const double t = own ? by_nb(tb[b]) : 0.0;
      tb[b] = t;
      if (own) a.T[static_cast<size_t>(b) * ldt + i] = t;
      ab[b] = t * ui;
    }
  }
  grid_sum<8>(ab, scratch, s_tot, x.work, 0, x, G);
  XSTAMP(5);
  if (tid == 0) {  // superweights to unit length (:880)
    double sq = 0.0;
    for (int b = 0; b < B; ++b) {
      s_a[b] = s_tot[b] / uu;
      sq = fma(s_a[b], s_a[b], sq);
    }
    sq = sqrt(sq);
    for (int b = 0; b < B; ++b) {
      s_a[b] /= sq;
      if (blockIdx.x == 0) a.a[b] = s_a[b];
    }
  }
  __syncthreads();
  // phase 2: superscore ts = T a, normalised (:882-883)
  double ts = 0.0;
  for (int b = 0; b < B; ++b) ts = fma(tb[b], s_a[b], ts);
  double one[1] = {ts * ts};
  grid_sum<1>(one, scratch, s_tot, x.work, 1, x, G);
  const double tsn = sqrt(s_tot[0]);
  __syncthreads();
  // phase 3 + 4: normalise, convergence metric against ts_old (:884-888), ts'ts, Y'ts (:899)
  const UniformDivisor by_tsn(tsn);
  const double t = own ? by_tsn(ts) : 0.0;
  const double d = own ? old - t : 0.0;
  if (own) {
    a.ts[i] = t;
    a.ts_old[i] = t;
  }
  {
    double dmax = warp_max(fabs(d));
    double dmin = warp_min(own ? fabs(d) : INFINITY);
    __syncthreads();
    if (lane == 0) {
      s_part[warp] = dmax;
      s_part[32 + warp] = dmin;
    }
    __syncthreads();
    if (tid == 0) {
      double mx = 0.0, mn = INFINITY;
      for (int wv = 0; wv < 32; ++wv) {
        mx = fmax(mx, s_part[wv]);
        mn = fmin(mn, s_part[32 + wv]);
      }
      double* slot = x.work + (static_cast<size_t>(5) * XMC_GMAX + blockIdx.x) * XMC_W;
      slot[0] = mx;
      slot[1] = mn;
    }
  }
  double v19[19];
  v19[0] = d * d;
  v19[1] = fabs(d);
  v19[2] = t * t;
#pragma unroll
  for (int c = 0; c < 16; ++c) v19[3 + c] = yv[c] * t;
  grid_sum<19>(v19, scratch, s_tot, x.work, 2, x, G);
  XSTAMP(6);
  const double tt = s_tot[2];
  if (tid == 0) {
    double mx = 0.0, mn = INFINITY;
    for (int g = 0; g < G; ++g) {
      const double* slot = x.work + (static_cast<size_t>(5) * XMC_GMAX + g) * XMC_W;
      mx = fmax(mx, __ldcg(slot));
      mn = fmin(mn, __ldcg(slot + 1));
    }
    double diff;
    switch (a.norm_kind) {
      case MBPLS_NORM_L1: diff = s_tot[1]; break;
      case MBPLS_NORM_MAX: diff = mx; break;
      case MBPLS_NORM_MIN: diff = mn; break;
      default: diff = sqrt(s_tot[0]);
    }
    s_ext[32] = diff;
    double sq = 0.0;
    for (int c = 0; c < q; ++c) {
      s_v[c] = s_tot[3 + c] / tt;
      sq = fma(s_v[c], s_v[c], sq);
      if (blockIdx.x == 0) a.v[c] = s_v[c];
    }
    s_ext[33] = sq;
  }
  __syncthreads();
  const double vv = s_ext[33];
  XSTAMP(7);
  // phase 5: Y scores u = Y v / v'v, normalised (:911-913)
  const UniformDivisor by_vv(vv);
  double un = 0.0;
  for (int c = 0; c < q; ++c) un = fma(yv[c], s_v[c], un);
  const double uval = own ? by_vv(un) : 0.0;
  double one2[1] = {uval * uval};
  grid_sum<1>(one2, scratch, s_tot, x.work, 3, x, G);
  const double unorm = sqrt(s_tot[0]);
  __syncthreads();
  const UniformDivisor by_un(unorm);
  const double unew = own ? by_un(uval) : 0.0;
  if (own) a.u[i] = unew;
  double one3[1] = {unew * unew};
  grid_sum<1>(one3, scratch, s_tot, x.work, 4, x, G);
  if (blockIdx.x == 0 && tid == 0) {
    int* ctrl = a.ctrl;
    const int trips = ctrl[MBPLS_CTRL_TRIPS] + 1;
    ctrl[MBPLS_CTRL_TRIPS] = trips;
    a.scal[MBPLS_SCAL_UU] = s_tot[0];
    a.scal[MBPLS_SCAL_TT] = tt;
    a.scal[MBPLS_SCAL_VV] = vv;
    if (trips > 1) {  // the first trip has nothing to compare with (:884-885)
      a.scal[MBPLS_SCAL_DIFF] = s_ext[32];
      if (!(s_ext[32] > a.max_tol)) ctrl[MBPLS_CTRL_DONE] = 1;
    }
    if (a.diff_trace && trips <= a.diff_trace_len) a.diff_trace[trips - 1] = trips > 1 ? s_ext[32] : 1.0;
  }
  XSTAMP(9);
}

// ------------------------------------------------------------------------------------------
// record_component: append the converged component to the result arrays (mbpls/mbpls.py:975-983).
// Component-major storage: Wt/W/P are K x p_local, Ts/U are K x ldt, T is B x K x ldt.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
record_component_kernel(mbpls_record_args a) {
  if (a.only_if_done && !*a.only_if_done) return;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int gsz = gridDim.x * blockDim.x;
  const double* red_nrm = a.red + static_cast<size_t>(a.nanmode ? 2 * a.B : a.B) * a.ldt;
  for (int j = gid; j < a.p; j += gsz) {
    const int b = block_of(a.block_off, a.B, j);
    const double wt = a.w[j];
    a.Wt_k[j] = wt;
    a.W_k[j] = wt / sqrt(red_nrm[b]);
  }
  for (int i = gid; i < a.n; i += gsz) {
    a.Ts_k[i] = a.ts[i];
    a.U_k[i] = a.u[i];
    for (int b = 0; b < a.B; ++b) a.T_k[static_cast<size_t>(b) * a.T_block_stride + i] = a.T[static_cast<size_t>(b) * a.ldt + i];
  }
  if (gid < a.q) a.V_k[gid] = a.v[gid];
  if (gid < a.B) a.A_k[gid] = a.a[gid] * a.a[gid];
}

// ------------------------------------------------------------------------------------------
// loadings + deflation on a resident feature (mbpls/mbpls.py:917-930, :968-969) and, optionally,
// the first w~ of the next component (:847/:856 with u = u0).
// ------------------------------------------------------------------------------------------
struct DeflateParams {
  int n;
  long ld;
  int nanmode;
  const double* ts;
  const double* u0;       // may be null (no fused next-XtU)
  const double* u0u0;     // device scalar u0'u0
  double* P_k;            // p_local loadings out
  double* w_next;         // p_local next w~ out (if u0)
  double* pss;            // p_local: p_j^2 (for explained variance), always written
};

template <bool CTA_WIDE>
__device__ __forceinline__ void deflate_resident(double* x, int j, const DeflateParams& P, double* scratch) {
  const int tid = CTA_WIDE ? threadIdx.x : (threadIdx.x & 31);
  const int nt = CTA_WIDE ? blockDim.x : 32;
  const int n = P.n;
  double v[3] = {0.0, 0.0, 0.0};  // num, masked ts'ts, NaN seen
  for (int i = tid; i < n; i += nt) {
    const double xi = x[i], t = P.ts[i];
    if (P.nanmode) {
      if (!isnan(xi)) {
        v[0] = fma(xi, t, v[0]);
        v[1] = fma(t, t, v[1]);
      } else {
        v[2] = 1.0;
      }
    } else {
      v[0] = fma(xi, t, v[0]);
    }
  }
  if (CTA_WIDE) block_sum<3>(v, scratch);
  else { v[0] = warp_sum(v[0]); v[1] = warp_sum(v[1]); v[2] = warp_sum(v[2]); }
  const double pj = (P.nanmode && v[2] > 0.0) ? v[0] / v[1] : v[0];
  double w[3] = {0.0, 0.0, 0.0};  // next: num, masked u0'u0, NaN seen
  for (int i = tid; i < n; i += nt) {
    const double xi = __dsub_rn(x[i], __dmul_rn(P.ts[i], pj));  // not contracted: the reference rounds ts*p first (:969)
    x[i] = xi;
    if (P.u0) {
      const double uv = P.u0[i];
      if (P.nanmode) {
        if (!isnan(xi)) {
          w[0] = fma(xi, uv, w[0]);
          w[1] = fma(uv, uv, w[1]);
        } else {
          w[2] = 1.0;
        }
      } else {
        w[0] = fma(xi, uv, w[0]);
      }
    }
  }
  if (P.u0) {
    if (CTA_WIDE) block_sum<3>(w, scratch);
    else { w[0] = warp_sum(w[0]); w[1] = warp_sum(w[1]); w[2] = warp_sum(w[2]); }
  }
  if (tid == 0) {
    P.P_k[j] = pj;
    P.pss[j] = pj * pj;
    if (P.u0) P.w_next[j] = (P.nanmode && w[2] > 0.0) ? w[0] / w[1] : w[0] / *P.u0u0;
  }
}

// Short features: one consumer warp per resident feature (8 consumer warps + the producer warp).
struct DeflateWarpOp {
  DeflateParams P;
  __device__ __forceinline__ void operator()(double* slab, int f0, int nf) {
    const int warp = threadIdx.x >> 5, nw = (blockDim.x >> 5) - 1;
    for (int f = warp; f < nf; f += nw) deflate_resident<false>(slab + static_cast<size_t>(f) * P.ld, f0 + f, P, nullptr);
  }
};

__global__ void __launch_bounds__(288) loadings_deflate_warp_kernel(double* __restrict__ Xt, StreamShape sh, DeflateParams P) {
  DeflateWarpOp op{P};
  stream_feature_slabs<true>(Xt, sh, op);
}

// Long features (n > 1024): one 1024-thread CTA (31 consumer warps + the producer warp) per resident
// feature.  Every consumer keeps its EPT samples of ts in registers for the whole kernel, so a feature
// costs two shared-memory reads and one write per element; u0 (fused next-XtU only) is streamed from
// L2 with EPT independent loads in flight.
template <int EPT>
struct DeflateWideOp {
  DeflateParams P;
  double* scratch;
  int nt;  // consumer threads
  double ts_r[EPT];
  __device__ __forceinline__ void init() {
    nt = blockDim.x - 32;
#pragma unroll
    for (int k = 0; k < EPT; ++k) {
      const int i = threadIdx.x + k * nt;
      ts_r[k] = (threadIdx.x < nt && i < P.n) ? P.ts[i] : 0.0;
    }
  }
  __device__ __forceinline__ void reduce3(double (&v)[3]) {
    if (P.nanmode) {
      block_sum_consumers<3>(v, scratch, nt);
    } else {
      double one[1] = {v[0]};
      block_sum_consumers<1>(one, scratch, nt);
      v[0] = one[0];
    }
  }
  __device__ __forceinline__ void operator()(double* slab, int f0, int nf) {
    for (int f = 0; f < nf; ++f) {
      double* x = slab + static_cast<size_t>(f) * P.ld;
      double v[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const int i = threadIdx.x + k * nt;
        const double xv = i < P.n ? x[i] : 0.0;
        if (P.nanmode) {
          if (!isnan(xv)) {
            v[0] = fma(xv, ts_r[k], v[0]);
            v[1] = fma(ts_r[k], ts_r[k], v[1]);
          } else {
            v[2] = 1.0;
          }
        } else {
          v[0] = fma(xv, ts_r[k], v[0]);
        }
      }
      reduce3(v);
      const double pj = (P.nanmode && v[2] > 0.0) ? v[0] / v[1] : v[0];
      double w[3] = {0.0, 0.0, 0.0};
      double u_r[EPT];
      if (P.u0) {
#pragma unroll
        for (int k = 0; k < EPT; ++k) {
          const int i = threadIdx.x + k * nt;
          u_r[k] = i < P.n ? __ldg(P.u0 + i) : 0.0;
        }
      }
#pragma unroll
      for (int k = 0; k < EPT; ++k) {
        const int i = threadIdx.x + k * nt;
        const double xi = __dsub_rn(i < P.n ? x[i] : 0.0, __dmul_rn(ts_r[k], pj));  // reference rounds ts*p first (:969)
        if (i < P.n) x[i] = xi;
        if (P.u0) {
          if (P.nanmode) {
            if (!isnan(xi)) {
              w[0] = fma(xi, u_r[k], w[0]);
              w[1] = fma(u_r[k], u_r[k], w[1]);
            } else {
              w[2] = 1.0;
            }
          } else {
            w[0] = fma(xi, u_r[k], w[0]);
          }
        }
      }
      if (P.u0) reduce3(w);
      if (threadIdx.x == 0) {
        P.P_k[f0 + f] = pj;
        P.pss[f0 + f] = pj * pj;
        if (P.u0) P.w_next[f0 + f] = (P.nanmode && w[2] > 0.0) ? w[0] / w[1] : w[0] / *P.u0u0;
      }
    }
  }
};

template <int EPT>
__global__ void __launch_bounds__(1024, 1) loadings_deflate_wide_kernel(double* __restrict__ Xt, StreamShape sh, DeflateParams P) {
  __shared__ double scratch[96];
  DeflateWideOp<EPT> op;
  op.P = P;
  op.scratch = scratch;
  op.init();
  stream_feature_slabs<true>(Xt, sh, op);
}

template <int EPT>
static void launch_deflate_wide(double* Xt, const StreamShape& sh, const DeflateParams& P, int grid, size_t smem, cudaStream_t st) {
  cudaFuncSetAttribute(loadings_deflate_wide_kernel<EPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  loadings_deflate_wide_kernel<EPT><<<grid, 1024, smem, st>>>(Xt, sh, P);
}

// Mid-length features (1024 < n <= 16384): REGISTER-resident variant.  A 256-thread CTA loads one whole
// feature into registers (EPT2 16-byte loads in flight per thread), reduces the loading, updates the
// registers and streams them back: exactly 1 read + 1 write of X, no shared-memory staging, and the
// n-vectors ts / u0 stay L1-resident.  Two CTAs per SM overlap one CTA's reductions with the other's
// loads (the single-CTA shared-memory pipeline above is latency-chain-bound for long features:
// measured 3.3 TB/s vs 6.0 TB/s for its copy skeleton, profiles/r1_notes.md).
template <int EPT2>
__global__ void __launch_bounds__(256, 2) loadings_deflate_regs_kernel(double* __restrict__ Xt, int p, DeflateParams P) {
  __shared__ double scratch[96];
  const int n2 = (P.n + 1) >> 1;  // 16-byte units; an odd tail pairs with the zero padding element
  const double2* __restrict__ ts2 = reinterpret_cast<const double2*>(P.ts);
  const double2* __restrict__ u2 = reinterpret_cast<const double2*>(P.u0);
  double2 xr[EPT2];
  int j = blockIdx.x;
  if (j < p) {
    const double2* x2 = reinterpret_cast<const double2*>(Xt + static_cast<size_t>(j) * P.ld);
#pragma unroll
    for (int k = 0; k < EPT2; ++k) {
      const int i = threadIdx.x + k * 256;
      xr[k] = i < n2 ? ld_stream(x2 + i) : make_double2(0.0, 0.0);
    }
  }
  while (j < p) {
    double v[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < EPT2; ++k) {
      const int i = threadIdx.x + k * 256;
      const double2 t = i < n2 ? ld_keep(ts2 + i) : make_double2(0.0, 0.0);
      if (P.nanmode) {
        if (!isnan(xr[k].x)) { v[0] = fma(xr[k].x, t.x, v[0]); v[1] = fma(t.x, t.x, v[1]); } else v[2] = 1.0;
        if (!isnan(xr[k].y)) { v[0] = fma(xr[k].y, t.y, v[0]); v[1] = fma(t.y, t.y, v[1]); } else v[2] = 1.0;
      } else {
        v[0] = fma(xr[k].x, t.x, v[0]);
        v[0] = fma(xr[k].y, t.y, v[0]);
      }
    }
    if (P.nanmode) {
      block_sum<3>(v, scratch);
    } else {
      double one[1] = {v[0]};
      block_sum<1>(one, scratch);
      v[0] = one[0];
    }
    const double pj = (P.nanmode && v[2] > 0.0) ? v[0] / v[1] : v[0];
    // update + store this feature
    const int jn = j + gridDim.x;
    double2* x2 = reinterpret_cast<double2*>(Xt + static_cast<size_t>(j) * P.ld);
    const double2* xn = reinterpret_cast<const double2*>(Xt + static_cast<size_t>(jn < p ? jn : j) * P.ld);
    double w[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < EPT2; ++k) {
      const int i = threadIdx.x + k * 256;
      if (i < n2) {
        const double2 t = ld_keep(ts2 + i);
        double2 xi;
        xi.x = __dsub_rn(xr[k].x, __dmul_rn(t.x, pj));  // the reference rounds ts*p before subtracting (:969)
        xi.y = __dsub_rn(xr[k].y, __dmul_rn(t.y, pj));
        st_stream(x2 + i, xi);
        if (P.u0) {
          const double2 uv = ld_keep(u2 + i);
          if (P.nanmode) {
            if (!isnan(xi.x)) { w[0] = fma(xi.x, uv.x, w[0]); w[1] = fma(uv.x, uv.x, w[1]); } else w[2] = 1.0;
            if (!isnan(xi.y)) { w[0] = fma(xi.y, uv.y, w[0]); w[1] = fma(uv.y, uv.y, w[1]); } else w[2] = 1.0;
          } else {
            w[0] = fma(xi.x, uv.x, w[0]);
            w[0] = fma(xi.y, uv.y, w[0]);
          }
        }
      }
    }
    // all registers are drained: issue the next feature's loads now, so they fly during the reduction below
    // (issued as one batch -- interleaving them with the L1 loads above serialises on the scoreboard)
    if (jn < p) {
#pragma unroll
      for (int k = 0; k < EPT2; ++k) {
        const int i = threadIdx.x + k * 256;
        if (i < n2) xr[k] = ld_stream(xn + i);
      }
    }
    if (P.u0) {
      if (P.nanmode) {
        block_sum<3>(w, scratch);
      } else {
        double one[1] = {w[0]};
        block_sum<1>(one, scratch);
        w[0] = one[0];
      }
    }
    if (threadIdx.x == 0) {
      P.P_k[j] = pj;
      P.pss[j] = pj * pj;
      if (P.u0) P.w_next[j] = (P.nanmode && w[2] > 0.0) ? w[0] / w[1] : w[0] / *P.u0u0;
    }
    j = jn;
  }
}

// Same algorithm with ts and u0 staged ONCE per CTA in shared memory: a 512-thread CTA runs two independent
// 256-thread feature pipelines (named barriers 1 and 2) that share the 2 x 8n-byte copies, so phase B no longer
// depends on L1 hit rates (ncu: 20 % of the ts/u0 sectors missed L1 in the 2-CTA variant and every miss sat
// on the update's critical path, profiles/r1_notes.md).
template <int EPT2>
__global__ void __launch_bounds__(512, 1) loadings_deflate_regs2_kernel(double* __restrict__ Xt, int p, DeflateParams P) {
  extern __shared__ __align__(16) double vecs[];  // ts[ld] | u0[ld]
  __shared__ double scratch2[2][24];
  const int n2 = (P.n + 1) >> 1;
  const int g = threadIdx.x >> 8, tid = threadIdx.x & 255;
  double* s_ts = vecs;
  double* s_u = vecs + P.ld;
  for (int i = threadIdx.x; i < P.ld; i += 512) {
    s_ts[i] = i < P.n ? P.ts[i] : 0.0;
    if (P.u0) s_u[i] = i < P.n ? P.u0[i] : 0.0;
  }
  __syncthreads();
  const double2* __restrict__ ts2 = reinterpret_cast<const double2*>(s_ts);
  const double2* __restrict__ u2 = reinterpret_cast<const double2*>(s_u);
  double* scratch = scratch2[g];
  const int stride = 2 * gridDim.x;
  double2 xr[EPT2];
  int j = 2 * blockIdx.x + g;
  if (j < p) {
    const double2* x2 = reinterpret_cast<const double2*>(Xt + static_cast<size_t>(j) * P.ld);
#pragma unroll
    for (int k = 0; k < EPT2; ++k) {
      const int i = tid + k * 256;
      xr[k] = i < n2 ? ld_stream(x2 + i) : make_double2(0.0, 0.0);
    }
  }
  while (j < p) {
    double v[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < EPT2; ++k) {
      const int i = tid + k * 256;
      const double2 t = i < n2 ? ts2[i] : make_double2(0.0, 0.0);
      if (P.nanmode) {
        if (!isnan(xr[k].x)) { v[0] = fma(xr[k].x, t.x, v[0]); v[1] = fma(t.x, t.x, v[1]); } else v[2] = 1.0;
        if (!isnan(xr[k].y)) { v[0] = fma(xr[k].y, t.y, v[0]); v[1] = fma(t.y, t.y, v[1]); } else v[2] = 1.0;
      } else {
        v[0] = fma(xr[k].x, t.x, v[0]);
        v[0] = fma(xr[k].y, t.y, v[0]);
      }
    }
    if (P.nanmode) {
      group_sum256<3>(v, scratch, g);
    } else {
      double one[1] = {v[0]};
      group_sum256<1>(one, scratch, g);
      v[0] = one[0];
    }
    const double pj = (P.nanmode && v[2] > 0.0) ? v[0] / v[1] : v[0];
    const int jn = j + stride;
    double2* x2 = reinterpret_cast<double2*>(Xt + static_cast<size_t>(j) * P.ld);
    const double2* xn = reinterpret_cast<const double2*>(Xt + static_cast<size_t>(jn < p ? jn : j) * P.ld);
    double w[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < EPT2; ++k) {
      const int i = tid + k * 256;
      if (i < n2) {
        const double2 t = ts2[i];
        double2 xi;
        xi.x = __dsub_rn(xr[k].x, __dmul_rn(t.x, pj));  // the reference rounds ts*p before subtracting (:969)
        xi.y = __dsub_rn(xr[k].y, __dmul_rn(t.y, pj));
        st_stream(x2 + i, xi);
        if (P.u0) {
          const double2 uv = u2[i];
          if (P.nanmode) {
            if (!isnan(xi.x)) { w[0] = fma(xi.x, uv.x, w[0]); w[1] = fma(uv.x, uv.x, w[1]); } else w[2] = 1.0;
            if (!isnan(xi.y)) { w[0] = fma(xi.y, uv.y, w[0]); w[1] = fma(uv.y, uv.y, w[1]); } else w[2] = 1.0;
          } else {
            w[0] = fma(xi.x, uv.x, w[0]);
            w[0] = fma(xi.y, uv.y, w[0]);
          }
        }
      }
    }
    if (jn < p) {  // registers are drained: the next feature's loads fly during the reduction below
#pragma unroll
      for (int k = 0; k < EPT2; ++k) {
        const int i = tid + k * 256;
        if (i < n2) xr[k] = ld_stream(xn + i);
      }
    }
    if (P.u0) {
      if (P.nanmode) {
        group_sum256<3>(w, scratch, g);
      } else {
        double one[1] = {w[0]};
        group_sum256<1>(one, scratch, g);
        w[0] = one[0];
      }
    }
    if (tid == 0) {
      P.P_k[j] = pj;
      P.pss[j] = pj * pj;
      if (P.u0) P.w_next[j] = (P.nanmode && w[2] > 0.0) ? w[0] / w[1] : w[0] / *P.u0u0;
    }
    j = jn;
  }
}

template <int EPT2>
static void launch_deflate_regs2(double* Xt, int p, const DeflateParams& P, cudaStream_t st) {
  int grid = num_sms();
  if (grid > (p + 1) / 2) grid = (p + 1) / 2;
  const size_t smem = 2 * static_cast<size_t>(P.ld) * sizeof(double);
  cudaFuncSetAttribute(loadings_deflate_regs2_kernel<EPT2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  loadings_deflate_regs2_kernel<EPT2><<<grid, 512, smem, st>>>(Xt, p, P);
}

template <int EPT2>
static void launch_deflate_regs(double* Xt, int p, const DeflateParams& P, cudaStream_t st) {
  int grid = num_sms() * 2;
  if (grid > p) grid = p;
  // ts and u0 (2 x 8n bytes) must stay L1-resident: give the unified L1/shared array entirely to L1
  cudaFuncSetAttribute(loadings_deflate_regs_kernel<EPT2>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
  loadings_deflate_regs_kernel<EPT2><<<grid, 256, 0, st>>>(Xt, p, P);
}

// global-memory fallback (feature too long for shared memory): 2 reads + 1 write
__global__ void __launch_bounds__(256) loadings_deflate_global_kernel(double* __restrict__ Xt, int p, DeflateParams P) {
  __shared__ double scratch[96];
  for (int j = blockIdx.x; j < p; j += gridDim.x) {
    deflate_resident<true>(Xt + static_cast<size_t>(j) * P.ld, j, P, scratch);
    __syncthreads();
  }
}

// w~ produced by the fused deflation has no per-CTA norm partials; this computes ||w~_b||^2 directly.
__global__ void __launch_bounds__(256)
block_sumsq_parts_kernel(const double* __restrict__ w, int p, int feats_per_cta, const int* __restrict__ block_off, int B,
                         double* __restrict__ norm_part, const int* __restrict__ done) {
  if (done && *done) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int f_begin = blockIdx.x * feats_per_cta, f_end = min(p, f_begin + feats_per_cta);
  for (int b = warp; b < B; b += nw) {
    const int lo = max(block_off[b], f_begin), hi = min(block_off[b + 1], f_end);
    double s = 0.0;
    for (int j = lo + lane; j < hi; j += 32) {
      const double v = w[j];
      s = fma(v, v, s);
    }
    s = warp_sum(s);
    if (lane == 0) norm_part[static_cast<size_t>(blockIdx.x) * B + b] = s;
  }
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

int mbpls_xtu_feats_per_cta(int p) {
  // ~32 CTAs per SM worth of work items keeps the tail short; multiple of 16 (8 warps x 2 features)
  long target = (static_cast<long>(p) + static_cast<long>(num_sms()) * 32 - 1) / (static_cast<long>(num_sms()) * 32);
  long f = ((target + 15) / 16) * 16;
  if (f < 16) f = 16;
  if (f > 4096) f = 4096;
  return static_cast<int>(f);
}

int mbpls_xw_ctas_per_sm(void) {
  static int v = -1;
  if (v < 0) {
    int b = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, xw_kernel<false>, 256, 0) != cudaSuccess || b < 1) b = 4;
    v = b;
  }
  return v;
}

int mbpls_xtu_num_ctas(int p) {
  const int f = mbpls_xtu_feats_per_cta(p);
  return (p + f - 1) / f;
}

int mbpls_nipals_xtu_f64(const double* Xt, long ld, int n, int p, const double* u, const double* uu, const int* block_off,
                         int B, double* w, double* norm_part, int nanmode, const int* done, void* stream) {
  if (!Xt || !u || !w || !block_off || B < 1 || (ld % 2) != 0) return MBPLS_ERR_ARG;
  if (p == 0) return MBPLS_OK;
  const int f = mbpls_xtu_feats_per_cta(p);
  const int grid = (p + f - 1) / f;
  const size_t smem = static_cast<size_t>(f) * sizeof(double);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (nanmode) xtu_kernel<true><<<grid, 256, smem, st>>>(Xt, ld, n, p, f, u, uu, block_off, B, w, norm_part, done);
  else xtu_kernel<false><<<grid, 256, smem, st>>>(Xt, ld, n, p, f, u, uu, block_off, B, w, norm_part, done);
  MBPLS_RETURN_LAST();
}

int mbpls_block_sumsq_parts_f64(const double* w, int p, const int* block_off, int B, double* norm_part, const int* done,
                                void* stream) {
  if (!w || !block_off || !norm_part) return MBPLS_ERR_ARG;
  if (p == 0) return MBPLS_OK;
  const int f = mbpls_xtu_feats_per_cta(p);
  const int grid = (p + f - 1) / f;
  block_sumsq_parts_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, p, f, block_off, B, norm_part, done);
  MBPLS_RETURN_LAST();
}

int mbpls_nipals_xw_f64(const double* Xt, long ld, int n, const double* w, const int* split_f0, const int* split_f1,
                        int nsplit, double* Tnum, double* Tden, long ldt, int nanmode, const int* done, void* stream) {
  if (!Xt || !w || !split_f0 || !split_f1 || !Tnum || (nanmode && !Tden) || (ld % 2) != 0 || (ldt % 2) != 0) return MBPLS_ERR_ARG;
  if (nsplit == 0 || n == 0) return MBPLS_OK;
  if (nsplit > 65535) return MBPLS_ERR_SIZE;
  dim3 grid((n + 511) / 512, nsplit);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (nanmode) xw_kernel<true><<<grid, 256, 0, st>>>(Xt, ld, n, w, split_f0, split_f1, Tnum, Tden, ldt, done);
  else xw_kernel<false><<<grid, 256, 0, st>>>(Xt, ld, n, w, split_f0, split_f1, Tnum, Tden, ldt, done);
  MBPLS_RETURN_LAST();
}

int mbpls_nipals_reduce_partials_f64(const double* Tnum, const double* Tden, long ldt, int n, int B,
                                     const int* block_split_off, const double* norm_part, int n_norm_parts, double* red,
                                     int nanmode, const int* done, void* stream) {
  if (!Tnum || !block_split_off || !norm_part || !red || B < 1 || B > 65535) return MBPLS_ERR_ARG;
  dim3 grid((n + 255) / 256, B);
  reduce_partials_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(Tnum, Tden, ldt, n, B, block_split_off, norm_part,
                                                                             n_norm_parts, red, nanmode, done);
  MBPLS_RETURN_LAST();
}

int mbpls_nipals_begin_component_f64(const double* u0, int n, double* u, double* scal, int* ctrl, void* stream) {
  if (!u0 || !u || !scal || !ctrl) return MBPLS_ERR_ARG;
  begin_component_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(u0, n, u, scal, ctrl);
  MBPLS_RETURN_LAST();
}

int mbpls_nipals_epilogue_f64(const mbpls_epilogue_args* args, void* stream) {
  if (!args || !args->red || !args->T || !args->u || !args->ts || !args->ts_old || !args->Yt || !args->a || !args->v ||
      !args->scal || !args->ctrl)
    return MBPLS_ERR_ARG;
  if (args->B < 1 || args->B > EPI_MAXB || args->q < 1 || args->q > EPI_MAXQ) return MBPLS_ERR_SIZE;
  if (args->nanmode && (!args->row_flag || !args->ycol_flag)) return MBPLS_ERR_ARG;
  nipals_epilogue_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(*args);
  MBPLS_RETURN_LAST();
}

int mbpls_nipals_xchg_epilogue_f64(const mbpls_xchg_args* x, int ctas, void* stream) {
  if (!x) return MBPLS_ERR_ARG;
  const mbpls_epilogue_args* args = &x->epi;
  if (!args->red || !args->T || !args->u || !args->ts || !args->ts_old || !args->Yt || !args->a || !args->v || !args->scal ||
      !args->ctrl || !x->Tnum || !x->block_split_off || !x->norm_part || !x->counters)
    return MBPLS_ERR_ARG;
  if (args->B < 1 || args->B > EPI_MAXB || args->q < 1 || args->q > EPI_MAXQ) return MBPLS_ERR_SIZE;
  if (args->nanmode && (!args->row_flag || !args->ycol_flag || !x->Tden)) return MBPLS_ERR_ARG;
  if (x->world < 1 || x->rank < 0 || x->rank >= x->world) return MBPLS_ERR_ARG;
  if (x->world > 1 && (!x->peer_bufs || x->seq == 0 || x->slot_elems < (args->nanmode ? 2 : 1) * args->B * args->ldt + args->B ||
                       x->flags_off < 2 * x->slot_elems))
    return MBPLS_ERR_ARG;
  // every CTA must be resident at once (B spins on flags): at most one CTA per SM
  int gmax = ctas > 0 ? ctas : 96;
  if (gmax > num_sms()) gmax = num_sms();
  const long items = static_cast<long>(args->nanmode ? 2 : 1) * args->B * args->n;
  mbpls_xchg_args xa = *x;
  int tpi = 1;  // warps per item: as many as keep the grid within gmax CTAs (a pure function of the shapes: reproducible sums)
  while (tpi < 32 && items * (2 * tpi) <= static_cast<long>(gmax) * 1024) tpi *= 2;
  xa.tpi = tpi;
  const long need = (items * tpi + 1023) / 1024;
  const int g = need < 1 ? 1 : (need > gmax ? gmax : static_cast<int>(need));
  const char* mc_env = getenv("MBPLS_XCHG_MC");  // read per call (tests switch it between fits)
  const bool mc_off = mc_env && atoi(mc_env) == 0;
  if (x->work && x->epoch > 0 && x->world == 1 && !mc_off && !args->nanmode && args->B <= 8 && args->q <= 16 && gmax <= XMC_GMAX && args->n >= 1) {
    // superlevel step on all CTAs: ch samples per CTA (a multiple of 32), one sample per thread
    int ch = (args->n + gmax - 1) / gmax;
    ch = (ch + 31) / 32 * 32;
    if (ch <= 1024) {
      const int gm = (args->n + ch - 1) / ch;
      int t2 = 1;
      while (t2 < 32 && static_cast<long>(args->B) * ch * (2 * t2) <= 1024) t2 *= 2;
      xa.tpi = t2;
      xa.ch = ch;
      xchg_epilogue_mc_kernel<<<gm, 1024, 0, static_cast<cudaStream_t>(stream)>>>(xa);
      MBPLS_RETURN_LAST();
    }
  }
  xchg_epilogue_kernel<<<g, 1024, 0, static_cast<cudaStream_t>(stream)>>>(xa);
  MBPLS_RETURN_LAST();
}

#ifdef MBPLS_XCHG_STAMPS
/* probe builds only (scripts/xchg_stamps.py): the phase timestamps of the most recent xchg_epilogue_kernel launch */
int mbpls_debug_xchg_stamps(unsigned long long* out16, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, g_xchg_stamps, 16 * sizeof(unsigned long long));
  if (reset) {
    unsigned long long zero[16] = {};
    cudaMemcpyToSymbol(g_xchg_stamps, zero, sizeof(zero));
  }
  return 0;
}
#endif

int mbpls_nipals_record_component_f64(const mbpls_record_args* args, void* stream) {
  if (!args) return MBPLS_ERR_ARG;
  const int m = args->p > args->n ? args->p : args->n;
  int grid = (m + 255) / 256;
  if (grid < 1) grid = 1;
  if (grid > num_sms() * 8) grid = num_sms() * 8;
  record_component_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*args);
  MBPLS_RETURN_LAST();
}

// mode: 0 = auto (resident pipeline when a feature fits in shared memory), 1 = global fallback
int mbpls_loadings_deflate_f64(double* Xt, long ld, int n, int p, const double* ts, const double* u0, const double* u0u0,
                               double* P_k, double* w_next, double* pss, int nanmode, int mode, void* stream) {
  if (!Xt || !ts || !P_k || !pss || (u0 && (!u0u0 || !w_next)) || ld < n || (ld % 16) != 0) return MBPLS_ERR_ARG;
  if (p == 0) return MBPLS_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeflateParams P{n, ld, nanmode, ts, u0, u0u0, P_k, w_next, pss};
  StreamShape sh;
  bool cta_wide = false;
  if ((mode == 0 || mode == 3) && n > 1024 && n <= 16384 &&
      2 * static_cast<size_t>(ld) * sizeof(double) + 1024 <= static_cast<size_t>(smem_optin())) {
    // register-resident features, ts/u0 shared-memory resident (two 256-thread pipelines per CTA)
    const int e = (((n + 1) >> 1) + 255) / 256;
    if (mode == 3) {  // experiment switch: the 2-CTA/SM variant that reads ts/u0 through L1
      if (e <= 20) launch_deflate_regs<20>(Xt, p, P, st);
      else launch_deflate_regs<32>(Xt, p, P, st);
    } else if (e <= 4) launch_deflate_regs2<4>(Xt, p, P, st);
    else if (e <= 8) launch_deflate_regs2<8>(Xt, p, P, st);
    else if (e <= 12) launch_deflate_regs2<12>(Xt, p, P, st);
    else if (e <= 16) launch_deflate_regs2<16>(Xt, p, P, st);
    else if (e <= 20) launch_deflate_regs2<20>(Xt, p, P, st);
    else if (e <= 24) launch_deflate_regs2<24>(Xt, p, P, st);
    else launch_deflate_regs2<32>(Xt, p, P, st);
  } else if (mode == 0 && n > 1024 && n <= 16384) {  // register-resident, vectors through L1 (n too long for smem copies)
    const int e = (((n + 1) >> 1) + 255) / 256;
    if (e <= 4) launch_deflate_regs<4>(Xt, p, P, st);
    else if (e <= 8) launch_deflate_regs<8>(Xt, p, P, st);
    else if (e <= 12) launch_deflate_regs<12>(Xt, p, P, st);
    else if (e <= 16) launch_deflate_regs<16>(Xt, p, P, st);
    else if (e <= 20) launch_deflate_regs<20>(Xt, p, P, st);
    else if (e <= 24) launch_deflate_regs<24>(Xt, p, P, st);
    else launch_deflate_regs<32>(Xt, p, P, st);
  } else if ((mode == 0 || mode == 2) && pick_stream_shape(ld, p, &sh, &cta_wide)) {  // mode 2: force the smem pipeline
    const size_t smem = stream_smem_bytes(sh);
    const int grid = stream_grid(sh, smem);
    if (cta_wide) {
      const int ept = (n + 991) / 992;  // 31 consumer warps
      if (ept <= 4) launch_deflate_wide<4>(Xt, sh, P, grid, smem, st);
      else if (ept <= 8) launch_deflate_wide<8>(Xt, sh, P, grid, smem, st);
      else if (ept <= 12) launch_deflate_wide<12>(Xt, sh, P, grid, smem, st);
      else launch_deflate_wide<16>(Xt, sh, P, grid, smem, st);
    } else {
      cudaFuncSetAttribute(loadings_deflate_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      loadings_deflate_warp_kernel<<<grid, 288, smem, st>>>(Xt, sh, P);
    }
  } else {
    int grid = p < num_sms() * 8 ? p : num_sms() * 8;
    loadings_deflate_global_kernel<<<grid, 256, 0, st>>>(Xt, p, P);
  }
  MBPLS_RETURN_LAST();
}

}  // extern "C"
