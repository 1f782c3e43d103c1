// One-pass NIPALS trip (mbpls/mbpls.py:843-875): block weights AND block scores from a single read of X.
//
// w~_j = x_j . u / u'u depends on feature j and u only, and the block score is a sum over features,
// t~_b = sum_j w~_j x_j (the division by ||w~_b|| is a scalar applied later, nipals.cu), so while feature j is
// resident on the SM both halves of the trip can be done: the reference's two GEMVs (X_b' u at :847/:856 and
// X_b w_b at :866/:875, i.e. two full reads of X per trip) become ONE read.  The price is state: a worker
// must hold a private n-vector of score accumulators next to the resident feature.
//
// Layout of one CTA (persistent, one per SM, 512 threads):
//   * the threads form G = 512/TG independent workers of TG threads; a worker owns one *split* (a contiguous
//     feature range inside one block) and walks its features one at a time;
//   * thread `tg` of a worker owns the same EPT 16-byte units (2 samples each) of every feature, so its
//     score accumulators live in registers for the whole kernel (2*EPT doubles) next to the current
//     feature's values (2*EPT doubles);
//   * a worker's features stream through its private ring of S shared-memory stages as 1-D bulk (TMA)
//     copies of one chunk (UC = TG*EPTC units, the last chunk of a feature shorter) each, completion
//     tracked by one mbarrier per stage.  A stage is released as soon as its chunk sits in registers:
//     every warp bumps a shared-memory counter after its reads and the LAST warp to arrive refills the
//     stage itself (expect_tx + bulk copy; a warp's reads of the stage have completed before the arithmetic
//     that precedes its arrival could issue, so the refill cannot overtake them).  No producer warp (a 17th warp caps the kernel at
//     96 registers and spills the accumulators), no polling, and the refill work lands on whichever warp
//     happens to be last instead of always delaying the same one (ncu on the first version, which had
//     one feeder thread per worker: 52 % of all stall samples at the worker barrier behind that warp);
//   * u lives in shared memory.
// Per feature: chunks -> registers with the dot product against u on the fly; one fixed-order reduction over
// the worker (warp shuffles + one named barrier); w~_j; then acc_i += w~_j x_ij in registers.  Partial
// scores go to Tnum[split][.] exactly like xw_kernel's, so reduce_partials / the epilogue are unchanged and
// results are bitwise reproducible (static split -> worker map, fixed summation order).
//
// NaN mode (:848-852, :867-872, :923-925): NaN entries are read as zero, so the masked numerators come out of the same
// arithmetic; every masked DENOMINATOR is derived from the NaN bit matrix (csrc/nanmask.cu) by two tiny kernels -- per
// feature before the pass (the reciprocal `rden[j]` the kernel multiplies with), per sample after it (Tden) -- instead
// of being accumulated here.  (The first NaN version kept per-sample "missing weight" sums in a second shared-memory
// n-vector per worker; that left a 64 KB ring and ran at 2.6-3.6 TB/s.)
//
// The second kernel closes a component the same way: loadings p_j = x_j . ts, rank-1 deflation
// x_j -= ts p_j (:917-930, :968-969) written back from registers, the next component's first weights
// w~_j = x_j(deflated) . u0 / u0'u0 (u restarts from the same Y column, :838) AND its first block-score
// partials, so the first trip of the next component needs no pass over X at all.
#include "launch.cuh"
#include "../../include/mbpls_b200.h"

using namespace mbpls;

namespace {

struct FusedArgs {
  const double* Xt;  // trip kernel: read only
  double* Xw;        // deflate kernel: same matrix, written in place
  long ld;
  int n;
  const double* u;   // trip: current Y scores; deflate: u0 (may be null -> no next-component outputs)
  const double* uu;  // device scalar u'u (resp. u0'u0)
  const double* ts;  // deflate only
  const double* rden;   // NaN mode: per-feature reciprocal masked denominator for u (trip) / ts (deflate)
  const double* rden2;  // NaN mode, deflate: same for u0
  const double* cmask;  // recurrence deflation, NaN mode: per-feature masked ts . u0
  const double* cscal;  // recurrence deflation: device scalar ts . u0
  double* gdef;         // recurrence deflation: running x_j(deflated) . u0 per feature (read, updated); trip: raw dot products out
  const int* split_f0;
  const int* split_f1;
  const int* split_block;
  int nsplit;
  int B;
  double* w;          // w~ out (trip) / next w~ out (deflate)
  double* norm_part;  // [nsplit * B], zero except (split, block of split)
  double* Tnum;       // [nsplit][ldt]
  long ldt;
  double* P_k;  // deflate: loadings out
  double* pss;  // deflate: p_j^2 out
  const int* done;
};

// TG threads per worker, EPTC units per thread per chunk, at most CPF chunks per feature, S ring stages
template <int TG, int EPTC_, int CPF_, int S_>
struct Cfg {
  static constexpr int kTG = TG, EPTC = EPTC_, CPF = CPF_, S = S_;
  static constexpr int G = 512 / TG, EPT = EPTC_ * CPF_, NW = TG / 32;
  static constexpr int UC = TG * EPTC_;  // units per (full) chunk
  static constexpr int MAX_UNITS = UC * CPF_;
};

// fixed-order sum over the TG threads of one worker; `buf` holds NV*NW doubles and must alternate between
// two buffers from call to call (a single named barrier per call is then enough).  After the barrier every
// thread reads all NW warp partials (broadcast loads) and adds them in the same fixed tree.
template <int NV, int TG>
__device__ __forceinline__ void worker_sum(double (&v)[NV], double* buf, int g, int warp_in_group, int lane) {
  constexpr int NW = TG / 32;
#pragma unroll
  for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) buf[k * NW + warp_in_group] = v[k];
  }
  asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(TG) : "memory");
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const double2* pr = reinterpret_cast<const double2*>(buf + k * NW);
    if (NW == 2) {
      const double2 p0 = pr[0];
      v[k] = p0.x + p0.y;
    } else {  // groups of four partials, then the groups: few live registers next to the accumulators
      double grp[NW >= 4 ? NW / 4 : 1];
#pragma unroll
      for (int q = 0; q < NW / 4; ++q) {
        const double2 p0 = pr[2 * q], p1 = pr[2 * q + 1];
        grp[q] = (p0.x + p0.y) + (p1.x + p1.y);
      }
      double t = grp[0];
      if (NW == 8) t = grp[0] + grp[1];
      if (NW == 16) t = (grp[0] + grp[1]) + (grp[2] + grp[3]);
      v[k] = t;
    }
  }
}

__device__ __forceinline__ unsigned atom_inc_smem(unsigned* p) {
  unsigned old;
  asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(p)) : "memory");
  return old;
}

// shared-memory carve-up shared by both kernels
template <class C>
struct Smem {
  double* vec0;     // u / ts
  double* vec1;     // miss (NaN trip, one per worker) / u0 (deflate)
  double* ring;     // [G][S][2*UC]
  uint64_t* full;   // [G][S] mbarriers: chunk landed
  unsigned* cnt;    // [G][S] (8-byte slots) warps that have finished reading the stage
  double* scratch;  // [G][2][3*NW]
  __device__ Smem(unsigned char* base, long ld, int nvec) {  // nvec n-vectors in front of the ring
    vec0 = reinterpret_cast<double*>(base);
    vec1 = vec0 + ld;
    ring = vec0 + static_cast<size_t>(nvec) * ld;
    full = reinterpret_cast<uint64_t*>(ring + static_cast<size_t>(C::G) * C::S * 2 * C::UC);
    cnt = reinterpret_cast<unsigned*>(full + C::G * C::S);
    scratch = reinterpret_cast<double*>(full + 2 * C::G * C::S);
  }
  __device__ __forceinline__ double* stage(int g, int s) const { return ring + (static_cast<size_t>(g) * C::S + s) * 2 * C::UC; }
};

template <class C>
size_t fused_smem_bytes(long ld, int nvec) {
  return static_cast<size_t>(nvec) * ld * 8 + static_cast<size_t>(C::G) * C::S * C::UC * 16 + static_cast<size_t>(C::G) * C::S * 16 +
         static_cast<size_t>(C::G) * 2 * 3 * C::NW * 8 + 64;
}

// chunk c of feature f of the matrix X -> ring stage (g, s)
template <class C>
__device__ __forceinline__ void issue_chunk(const double* __restrict__ X, long ld, int units, int g, int s, int f, int c,
                                            const Smem<C>& sm) {
  const int nun = min(C::UC, units - c * C::UC);
  const uint32_t bytes = static_cast<uint32_t>(nun) * 16u;
  uint64_t* fb = &sm.full[g * C::S + s];
  mbar_arrive_expect_tx(fb, bytes);
  bulk_g2s(sm.stage(g, s), X + static_cast<size_t>(f) * ld + static_cast<size_t>(c) * C::UC * 2, bytes, fb);
}

template <class C>
__device__ __forceinline__ void init_sync(const Smem<C>& sm) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < C::G * C::S; ++i) {
      mbar_init(&sm.full[i], 1);
      sm.cnt[2 * i] = 0;
    }
    fence_barrier_init();
  }
}

// first S chunks of a worker (its thread 0, after the CTA-wide barrier that follows init_sync)
template <class C>
__device__ __forceinline__ void prime_ring(const double* __restrict__ X, long ld, int units, int ncf, int g, int f0, int f1,
                                           const Smem<C>& sm) {
  int f = f0, c = 0;
  for (int s = 0; s < C::S && f < f1; ++s) {
    issue_chunk<C>(X, ld, units, g, s, f, c, sm);
    if (++c == ncf) { c = 0; ++f; }
  }
}

// The calling warp is done reading stage (g, s).  Returns the number of warps that had said so before (lane 0 only).
// Relaxed atomic: every shared-memory read of the stage feeds a dot-product FMA that precedes this call in program order,
// and an instruction cannot issue before its operands have arrived, so the reads have completed when the counter moves.
template <class C>
__device__ __forceinline__ unsigned arrive_stage(int g, int s, int lane, const Smem<C>& sm) {
  __syncwarp();
  return lane == 0 ? atom_inc_smem(&sm.cnt[2 * (g * C::S + s)]) : 0u;
}

// The last warp of the worker to arrive on a stage (which held chunk c of feature j) refills it with the chunk S
// positions further down the worker's stream, immediately (a free stage is lost ring depth).  -DMBPLS_FUSED_PROXY_FENCE adds
// an explicit generic->async proxy fence in front of the bulk copy.
template <class C>
__device__ __forceinline__ void refill_if_last(unsigned old, const double* __restrict__ X, long ld, int units, int ncf, int g, int s,
                                               int j, int c, int f1, int lane, const Smem<C>& sm) {
  if (lane == 0 && old == C::NW - 1) {
    *reinterpret_cast<volatile unsigned*>(&sm.cnt[2 * (g * C::S + s)]) = 0;
    int c2 = c + C::S, f2 = j;
    while (c2 >= ncf) { c2 -= ncf; ++f2; }
    if (f2 < f1) {
#ifdef MBPLS_FUSED_PROXY_FENCE
      fence_proxy_async_smem();
#endif
      issue_chunk<C>(X, ld, units, g, s, f2, c2, sm);
    }
  }
}

// ------------------------------------------------------------------------------------------
// one NIPALS trip in one pass
//
// Software pipeline without extra registers: iteration j first applies the PREVIOUS feature's weight to the
// accumulators (acc += w~_{j-1} x) unit by unit and reloads each x register with feature j right behind it,
// so the score update hides under the shared-memory-bound load phase.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 nan_to_zero(double2 v) {
  if (isnan(v.x)) v.x = 0.0;
  if (isnan(v.y)) v.y = 0.0;
  return v;
}

template <bool NANMODE, class C>
__global__ void __launch_bounds__(512, 1) fused_trip_kernel(const FusedArgs a) {
  if (a.done && *a.done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const long ld = a.ld;
  const int units = static_cast<int>(ld >> 1);
  const int ncf = (units + C::UC - 1) / C::UC;
  const Smem<C> sm(smem_raw, ld, 1);  // u
  init_sync<C>(sm);
  for (int i = threadIdx.x; i < ld; i += blockDim.x) sm.vec0[i] = i < a.n ? a.u[i] : 0.0;
  __syncthreads();

  const int g = threadIdx.x / C::kTG, tg = threadIdx.x % C::kTG;
  const int lane = threadIdx.x & 31, wig = tg >> 5;
  const int wk = blockIdx.x * C::G + g;
  if (wk >= a.nsplit) return;
  const int f0 = a.split_f0[wk], f1 = a.split_f1[wk];
  if (tg == 0) prime_ring<C>(a.Xt, ld, units, ncf, g, f0, f1, sm);
  const double inv_uu = 1.0 / *a.uu;
  const double2* __restrict__ u2 = reinterpret_cast<const double2*>(sm.vec0);
  double* scratch = sm.scratch + static_cast<size_t>(g) * 2 * 3 * C::NW;

  double2 acc[C::EPT], x[C::EPT];
#pragma unroll
  for (int k = 0; k < C::EPT; ++k) acc[k] = x[k] = make_double2(0.0, 0.0);
  double normsq = 0.0, wj = 0.0;
  int s = 0;
  uint32_t ph = 0;
  int flip = 0;

  for (int j = f0; j <= f1; ++j) {
    const bool load = j < f1;  // the last iteration only applies the last weight
    const double rd = (NANMODE && load) ? a.rden[j] : inv_uu;  // 1 / (masked) u'u of this feature
    double numa = 0.0, numb = 0.0, numc = 0.0, numd = 0.0;
    int s_use = s;  // s: next stage to load from; s_use: stage of the chunk being consumed

    // previous feature's weight into the accumulators, then this feature's chunk c into the freed registers
    auto load_chunk = [&](const int c) {
      const bool have = load && c < ncf;
      const double2* __restrict__ xs = reinterpret_cast<const double2*>(sm.stage(g, s));
      if (have) {
        mbar_wait(&sm.full[g * C::S + s], ph);
        if (++s == C::S) { s = 0; ph ^= 1u; }
      }
      const bool whole = c + 1 < ncf;  // every chunk but the last of a feature is full: no bounds checks
#pragma unroll
      for (int e = 0; e < C::EPTC; ++e) {
        const int l = tg + e * C::kTG, gi = c * C::UC + l, k = c * C::EPTC + e;
        acc[k].x = fma(wj, x[k].x, acc[k].x);
        acc[k].y = fma(wj, x[k].y, acc[k].y);
        double2 xv = make_double2(0.0, 0.0), uv = make_double2(0.0, 0.0);
        if (have && (whole || gi < units)) {
          xv = xs[l];
          uv = u2[gi];
          if (NANMODE) xv = nan_to_zero(xv);  // masked sums: a missing entry contributes nothing (:848-852, :867-872)
        }
        if (have) {
          if (e & 1) {
            numc = fma(xv.x, uv.x, numc);
            numd = fma(xv.y, uv.y, numd);
          } else {
            numa = fma(xv.x, uv.x, numa);
            numb = fma(xv.y, uv.y, numb);
          }
        }
        x[k] = xv;
      }
    };
    // hand the stage back as soon as the chunk sits in registers; refill at once: every cycle a free stage sits idle is
    // ring depth lost (deferring the check by one chunk to hide the atomic's latency cost 10 % on the shallow rings, and
    // issuing chunk c+1's shared-memory loads before consuming chunk c changed nothing: profiles/r1_notes.md)
    auto release_chunk = [&](const int c) {
      if (!(load && c < ncf)) return;
      refill_if_last<C>(arrive_stage<C>(g, s_use, lane, sm), a.Xt, ld, units, ncf, g, s_use, j, c, f1, lane, sm);
      if (++s_use == C::S) s_use = 0;
    };
#pragma unroll
    for (int c = 0; c < C::CPF; ++c) {
      load_chunk(c);
      release_chunk(c);
    }
    if (!load) break;
    double one[1] = {(numa + numb) + (numc + numd)};
    worker_sum<1, C::kTG>(one, scratch + flip * 3 * C::NW, g, wig, lane);
    flip ^= 1;
    wj = one[0] * rd;
    if (tg == 0) {
      a.w[j] = wj;
      if (a.gdef) a.gdef[j] = one[0];  // x_j . u (numerator only): seeds the running x_j . u0 of the recurrence deflation
    }
    normsq = fma(wj, wj, normsq);
  }

  // partial block scores of this split (the layout xw_kernel writes)
  double2* tn = reinterpret_cast<double2*>(a.Tnum + static_cast<size_t>(wk) * a.ldt);
#pragma unroll
  for (int k = 0; k < C::EPT; ++k) {
    const int gi = (k / C::EPTC) * C::UC + tg + (k % C::EPTC) * C::kTG;
    if (gi < units) tn[gi] = acc[k];
  }
  if (tg == 0) a.norm_part[static_cast<size_t>(wk) * a.B + a.split_block[wk]] = normsq;
}

// ------------------------------------------------------------------------------------------
// loadings + deflation + the whole first trip of the next component; same pipeline: the score update of
// feature j-1 rides on the load phase of feature j.  NaN mode: NaN entries take part as zeros and are written
// back as NaN (:969 keeps them); loadings and next weights use the masked reciprocal denominators rden / rden2.
// ------------------------------------------------------------------------------------------
template <bool NANMODE, class C>
__global__ void __launch_bounds__(512, 1) fused_deflate_kernel(const FusedArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const long ld = a.ld;
  const int units = static_cast<int>(ld >> 1);
  const int ncf = (units + C::UC - 1) / C::UC;
  const Smem<C> sm(smem_raw, ld, 2);  // ts | u0
  init_sync<C>(sm);
  const bool next = a.u != nullptr;
  for (int i = threadIdx.x; i < ld; i += blockDim.x) {
    sm.vec0[i] = i < a.n ? a.ts[i] : 0.0;
    sm.vec1[i] = (next && i < a.n) ? a.u[i] : 0.0;
  }
  __syncthreads();

  const int g = threadIdx.x / C::kTG, tg = threadIdx.x % C::kTG;
  const int lane = threadIdx.x & 31, wig = tg >> 5;
  const int wk = blockIdx.x * C::G + g;
  if (wk >= a.nsplit) return;
  const int f0 = a.split_f0[wk], f1 = a.split_f1[wk];
  if (tg == 0) prime_ring<C>(a.Xw, ld, units, ncf, g, f0, f1, sm);
  const double inv_uu = next ? 1.0 / *a.uu : 1.0;
  const double2* __restrict__ ts2 = reinterpret_cast<const double2*>(sm.vec0);
  const double2* __restrict__ u2 = reinterpret_cast<const double2*>(sm.vec1);
  double* scratch = sm.scratch + static_cast<size_t>(g) * 2 * 3 * C::NW;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

  double2 acc[C::EPT], x[C::EPT];
#pragma unroll
  for (int k = 0; k < C::EPT; ++k) acc[k] = x[k] = make_double2(0.0, 0.0);
  double normsq = 0.0, wj = 0.0;
  int s = 0;
  uint32_t ph = 0;
  int flip = 0;

  for (int j = f0; j <= f1; ++j) {
    const bool load = j < f1;
    const double rdp = (NANMODE && load) ? a.rden[j] : 1.0;                      // loadings: 1 / masked ts'ts (dense: not divided, :920)
    const double rdw = (NANMODE && load && next) ? a.rden2[j] : inv_uu;          // next weights: 1 / masked u0'u0
    double pa = 0.0, pb = 0.0, pc = 0.0, pd = 0.0;
    uint32_t mx = 0, my = 0;  // NaN mode: which of this thread's entries are NaN (they must be stored back as NaN)
    int s_use = s;

    auto load_chunk = [&](const int c) {
      const bool have = load && c < ncf;
      const double2* __restrict__ xs = reinterpret_cast<const double2*>(sm.stage(g, s));
      if (have) {
        mbar_wait(&sm.full[g * C::S + s], ph);
        if (++s == C::S) { s = 0; ph ^= 1u; }
      }
      const bool whole = c + 1 < ncf;
#pragma unroll
      for (int e = 0; e < C::EPTC; ++e) {
        const int l = tg + e * C::kTG, gi = c * C::UC + l, k = c * C::EPTC + e;
        acc[k].x = fma(wj, x[k].x, acc[k].x);  // next component's block-score partials with the previous feature's weight
        acc[k].y = fma(wj, x[k].y, acc[k].y);
        double2 xv = make_double2(0.0, 0.0), tv = make_double2(0.0, 0.0);
        if (have && (whole || gi < units)) {
          xv = xs[l];
          tv = ts2[gi];
          if (NANMODE) {
            if (isnan(xv.x)) { xv.x = 0.0; mx |= 1u << k; }
            if (isnan(xv.y)) { xv.y = 0.0; my |= 1u << k; }
          }
        }
        if (e & 1) {
          pc = fma(xv.x, tv.x, pc);
          pd = fma(xv.y, tv.y, pd);
        } else {
          pa = fma(xv.x, tv.x, pa);
          pb = fma(xv.y, tv.y, pb);
        }
        x[k] = xv;
      }
    };
    auto release_chunk = [&](const int c) {
      if (!(load && c < ncf)) return;
      refill_if_last<C>(arrive_stage<C>(g, s_use, lane, sm), a.Xw, ld, units, ncf, g, s_use, j, c, f1, lane, sm);
      if (++s_use == C::S) s_use = 0;
    };
#pragma unroll
    for (int c = 0; c < C::CPF; ++c) {
      load_chunk(c);
      release_chunk(c);
    }
    if (!load) break;
    double one[1] = {(pa + pb) + (pc + pd)};
    worker_sum<1, C::kTG>(one, scratch + flip * 3 * C::NW, g, wig, lane);
    flip ^= 1;
    const double pj = one[0] * rdp;
    double2* __restrict__ xg = reinterpret_cast<double2*>(a.Xw + static_cast<size_t>(j) * ld);
    double wa = 0.0, wb = 0.0, wc = 0.0, wd = 0.0;
#pragma unroll
    for (int k = 0; k < C::EPT; ++k) {
      const int gi = (k / C::EPTC) * C::UC + tg + (k % C::EPTC) * C::kTG;
      if (((k / C::EPTC) + 1 < ncf) || gi < units) {
        const double2 tv = ts2[gi];
        double2 xn;
        xn.x = __dsub_rn(x[k].x, __dmul_rn(tv.x, pj));  // the reference rounds ts*p before subtracting (:969)
        xn.y = __dsub_rn(x[k].y, __dmul_rn(tv.y, pj));
        if (NANMODE) {
          double2 out = xn;
          if ((mx >> k) & 1u) { out.x = qnan; xn.x = 0.0; }
          if ((my >> k) & 1u) { out.y = qnan; xn.y = 0.0; }
          st_stream(xg + gi, out);
        } else {
          st_stream(xg + gi, xn);
        }
        x[k] = xn;
        if (next) {
          const double2 uv = u2[gi];
          if (k & 1) {
            wc = fma(xn.x, uv.x, wc);
            wd = fma(xn.y, uv.y, wd);
          } else {
            wa = fma(xn.x, uv.x, wa);
            wb = fma(xn.y, uv.y, wb);
          }
        }
      }
    }
    if (next) {
      double two[1] = {(wa + wb) + (wc + wd)};
      worker_sum<1, C::kTG>(two, scratch + flip * 3 * C::NW, g, wig, lane);
      flip ^= 1;
      wj = two[0] * rdw;
      normsq = fma(wj, wj, normsq);
    }
    if (tg == 0) {
      a.P_k[j] = pj;
      a.pss[j] = pj * pj;
      if (next) a.w[j] = wj;
    }
  }
  if (!next) return;
  double2* tn = reinterpret_cast<double2*>(a.Tnum + static_cast<size_t>(wk) * a.ldt);
#pragma unroll
  for (int k = 0; k < C::EPT; ++k) {
    const int gi = (k / C::EPTC) * C::UC + tg + (k % C::EPTC) * C::kTG;
    if (gi < units) tn[gi] = acc[k];
  }
  if (tg == 0) a.norm_part[static_cast<size_t>(wk) * a.B + a.split_block[wk]] = normsq;
}

// ------------------------------------------------------------------------------------------
// Recurrence deflation (opt-in, set_runtime(deflate_rec=True)): the same pass without a resident u0.
//
// With p_j = x_j . ts in hand, the dot product of the DEFLATED feature with u0 follows from that of the undeflated one:
// x_j(deflated) . u0 = x_j . u0 - p_j (ts . u0)   (masked data: all sums over the observed samples of the feature).
// So a_j = x_j . u0 obeys a recurrence over the components.  Keeping that scalar per feature in global memory (`gdef`,
// seeded with the raw dot products of a first trip, whose u IS u0, and re-seeded the same way every few components so that
// rounding cannot pile up) removes u0 from shared memory: the kernel then has the trip kernel's footprint -- one n-vector
// and the ring -- i.e. 40 KB chunks instead of 16 KB at n = 10,000 (30.4 -> 26.6 ms for 160 GB, 6.0 TB/s read + write),
// and the next weight is known right after the ONE reduction, so update, write-back and score accumulation of feature j-1
// ride, unit by unit, on the load loop of feature j (each register is updated, stored, accumulated and reloaded in turn).
// Not the default: the carried scalar is noisier than a fresh dot product by the ratio |x_j| / |x_j deflated|, which can
// push diff_t of a late component's second trip over the reference's max_tol = 1e-14 and cost a third trip
// (profiles/r1_notes.md).  A variant that keeps u0 resident and forms x_j . u0 afresh (one reduction, same footprint as
// fused_deflate_kernel) measured 29.0 vs 30.0 ms and was dropped.
// ------------------------------------------------------------------------------------------
template <bool NANMODE, class C>
__global__ void __launch_bounds__(512, 1) fused_deflate3_kernel(const FusedArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const long ld = a.ld;
  const int units = static_cast<int>(ld >> 1);
  const int ncf = (units + C::UC - 1) / C::UC;
  const Smem<C> sm(smem_raw, ld, 1);  // ts
  init_sync<C>(sm);
  const bool next = a.gdef != nullptr;
  for (int i = threadIdx.x; i < ld; i += blockDim.x) sm.vec0[i] = i < a.n ? a.ts[i] : 0.0;
  const double c_dense = next ? *a.cscal : 0.0;  // ts . u0
  __syncthreads();

  const int g = threadIdx.x / C::kTG, tg = threadIdx.x % C::kTG;
  const int lane = threadIdx.x & 31, wig = tg >> 5;
  const int wk = blockIdx.x * C::G + g;
  if (wk >= a.nsplit) return;
  const int f0 = a.split_f0[wk], f1 = a.split_f1[wk];
  if (tg == 0) prime_ring<C>(a.Xw, ld, units, ncf, g, f0, f1, sm);
  const double inv_uu = next ? 1.0 / *a.uu : 1.0;
  const double2* __restrict__ ts2 = reinterpret_cast<const double2*>(sm.vec0);
  double* scratch = sm.scratch + static_cast<size_t>(g) * 2 * 3 * C::NW;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);

  double2 acc[C::EPT], x[C::EPT];
#pragma unroll
  for (int k = 0; k < C::EPT; ++k) acc[k] = x[k] = make_double2(0.0, 0.0);
  double normsq = 0.0, wj = 0.0, pj = 0.0;
  uint32_t mx = 0, my = 0;  // NaN mode: NaN positions of the feature held in x (written back as NaN)
  int s = 0;
  uint32_t ph = 0;
  int flip = 0;

  for (int j = f0; j <= f1; ++j) {
    const bool load = j < f1, store = j > f0;
    const double rdp = (NANMODE && load) ? a.rden[j] : 1.0;               // loadings: 1 / masked ts'ts (dense: not divided, :920)
    const double rdw = (NANMODE && load && next) ? a.rden2[j] : inv_uu;   // next weights: 1 / masked u0'u0
    const double cj = (NANMODE && load && next) ? a.cmask[j] : c_dense;   // (masked) ts . u0
    double2* __restrict__ xg = reinterpret_cast<double2*>(a.Xw + static_cast<size_t>(j - 1) * ld);  // row of the feature held in x
    const double gj = (load && next) ? a.gdef[j] : 0.0;  // x_j . u0 before this deflation
    double pa = 0.0, pb = 0.0, pc = 0.0, pd = 0.0;
    uint32_t nmx = 0, nmy = 0;
#pragma unroll
    for (int c = 0; c < C::CPF; ++c) {
      const bool have = load && c < ncf;
      const double2* __restrict__ xs = reinterpret_cast<const double2*>(sm.stage(g, s));
      if (have) mbar_wait(&sm.full[g * C::S + s], ph);
      const bool whole = c + 1 < ncf;
#pragma unroll
      for (int e = 0; e < C::EPTC; ++e) {
        const int l = tg + e * C::kTG, gi = c * C::UC + l, k = c * C::EPTC + e;
        if (c < ncf && (whole || gi < units)) {
          const double2 tv = ts2[gi];
          if (store) {  // feature j-1: deflate, write back, accumulate the next component's score partials
            double2 xn;
            xn.x = __dsub_rn(x[k].x, __dmul_rn(tv.x, pj));  // the reference rounds ts*p before subtracting (:969)
            xn.y = __dsub_rn(x[k].y, __dmul_rn(tv.y, pj));
            if (NANMODE) {
              double2 out = xn;
              if ((mx >> k) & 1u) { out.x = qnan; xn.x = 0.0; }
              if ((my >> k) & 1u) { out.y = qnan; xn.y = 0.0; }
              st_stream(xg + gi, out);
            } else {
              st_stream(xg + gi, xn);
            }
            acc[k].x = fma(wj, xn.x, acc[k].x);
            acc[k].y = fma(wj, xn.y, acc[k].y);
          }
          if (have) {  // feature j: into the freed registers, both dot products on the fly
            double2 xv = xs[l];
            if (NANMODE) {
              if (isnan(xv.x)) { xv.x = 0.0; nmx |= 1u << k; }
              if (isnan(xv.y)) { xv.y = 0.0; nmy |= 1u << k; }
            }
            if (e & 1) {
              pc = fma(xv.x, tv.x, pc);
              pd = fma(xv.y, tv.y, pd);
            } else {
              pa = fma(xv.x, tv.x, pa);
              pb = fma(xv.y, tv.y, pb);
            }
            x[k] = xv;
          }
        }
      }
      if (have) {
        refill_if_last<C>(arrive_stage<C>(g, s, lane, sm), a.Xw, ld, units, ncf, g, s, j, c, f1, lane, sm);
        if (++s == C::S) { s = 0; ph ^= 1u; }
      }
    }
    if (!load) break;
    mx = nmx;
    my = nmy;
    double v[1] = {(pa + pb) + (pc + pd)};
    worker_sum<1, C::kTG>(v, scratch + flip * 3 * C::NW, g, wig, lane);
    flip ^= 1;
    pj = v[0] * rdp;
    const double gnew = gj - pj * cj;  // x_j(deflated) . u0
    if (next) {
      wj = gnew * rdw;
      normsq = fma(wj, wj, normsq);
    }
    if (tg == 0) {
      a.P_k[j] = pj;
      a.pss[j] = pj * pj;
      if (next) {
        a.w[j] = wj;
        a.gdef[j] = gnew;
      }
    }
  }
  if (!next) return;
  double2* tn = reinterpret_cast<double2*>(a.Tnum + static_cast<size_t>(wk) * a.ldt);
#pragma unroll
  for (int k = 0; k < C::EPT; ++k) {
    const int gi = (k / C::EPTC) * C::UC + tg + (k % C::EPTC) * C::kTG;
    if (gi < units) tn[gi] = acc[k];
  }
  if (tg == 0) a.norm_part[static_cast<size_t>(wk) * a.B + a.split_block[wk]] = normsq;
}

// Configurations by feature length (units = ld/2 16-byte units per feature <= TG*EPTC*CPF).  Measured (profiles/r1_notes.md):
// every chunk costs a worker ~0.2 us of handshakes (wait, counter, refill), so chunks are as large as the ring allows:
// the 16 KB x 8 ring ran the n = 10,000 trip at 4.9 TB/s, the 40 KB x 3 ring runs it at 6.6 TB/s.
using CfgA = Cfg<512, 5, 2, 3>;   // ld <= 10240: one worker per CTA, 40 KB chunks (u + ring = 200 KB)
using CfgA4 = Cfg<512, 2, 5, 4>;  // same length with two resident n-vectors (deflate: ts and u0): only 64 KB of ring left
using CfgB = Cfg<256, 5, 2, 3>;   // ld <= 5120: two workers, 20 KB chunks
using CfgC = Cfg<128, 5, 2, 4>;   // ld <= 2560: four workers, 10 KB chunks
using CfgD = Cfg<64, 10, 1, 2>;   // ld <= 1280: eight workers, one chunk per feature

template <bool NANMODE, class C>
int launch_trip(const FusedArgs& a, cudaStream_t st) {
  const size_t smem = fused_smem_bytes<C>(a.ld, 1);
  if (smem > static_cast<size_t>(smem_optin()) || (a.ld >> 1) > C::MAX_UNITS) return MBPLS_ERR_SIZE;
  cudaFuncSetAttribute(fused_trip_kernel<NANMODE, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  const int grid = (a.nsplit + C::G - 1) / C::G;
  fused_trip_kernel<NANMODE, C><<<grid, 512, smem, st>>>(a);
  return MBPLS_OK;
}

template <bool NANMODE, class C>
int launch_deflate(const FusedArgs& a, cudaStream_t st) {
  const size_t smem = fused_smem_bytes<C>(a.ld, 2);
  if (smem > static_cast<size_t>(smem_optin()) || (a.ld >> 1) > C::MAX_UNITS) return MBPLS_ERR_SIZE;
  const int grid = (a.nsplit + C::G - 1) / C::G;
  cudaFuncSetAttribute(fused_deflate_kernel<NANMODE, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  fused_deflate_kernel<NANMODE, C><<<grid, 512, smem, st>>>(a);
  return MBPLS_OK;
}

template <bool NANMODE, class C>
int launch_deflate3(const FusedArgs& a, cudaStream_t st) {
  const size_t smem = fused_smem_bytes<C>(a.ld, 1);
  if (smem > static_cast<size_t>(smem_optin()) || (a.ld >> 1) > C::MAX_UNITS) return MBPLS_ERR_SIZE;
  const int grid = (a.nsplit + C::G - 1) / C::G;
  cudaFuncSetAttribute(fused_deflate3_kernel<NANMODE, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  fused_deflate3_kernel<NANMODE, C><<<grid, 512, smem, st>>>(a);
  return MBPLS_OK;
}

int config_of(long ld) {  // 0: unsupported
  if (ld < 16 || (ld % 16) != 0) return 0;
  const long units = ld >> 1;
  if (units <= 640) return 4;
  if (units <= 1280) return 3;
  if (units <= 2560) return 2;
  if (units <= 5120) return 1;
  return 0;
}

}  // namespace

extern "C" {

/* workers (splits) per CTA of the one-pass kernels for this leading dimension; 0 = feature too long */
int mbpls_fused_workers_per_cta(long ld) {
  switch (config_of(ld)) {
    case 1: return CfgA::G;
    case 2: return CfgB::G;
    case 3: return CfgC::G;
    case 4: return CfgD::G;
    default: return 0;
  }
}

int mbpls_nipals_fused_trip_f64(const double* Xt, long ld, int n, const double* u, const double* uu, const double* rden,
                                const int* split_f0, const int* split_f1, const int* split_block, int nsplit, int B, double* w,
                                double* norm_part, double* Tnum, long ldt, double* dots_out, const int* done, void* stream) {
  if (!Xt || !u || !uu || !split_f0 || !split_f1 || !split_block || !w || !norm_part || !Tnum || ld < n || ldt < ld || B < 1)
    return MBPLS_ERR_ARG;
  if (nsplit == 0) return MBPLS_OK;
  FusedArgs a{Xt, nullptr, ld, n, u, uu, nullptr, rden, nullptr, nullptr, nullptr, dots_out, split_f0, split_f1, split_block, nsplit, B,
              w, norm_part, Tnum, ldt, nullptr, nullptr, done};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = MBPLS_ERR_SIZE;
  switch (config_of(ld)) {
    case 1: rc = rden ? launch_trip<true, CfgA>(a, st) : launch_trip<false, CfgA>(a, st); break;
    case 2: rc = rden ? launch_trip<true, CfgB>(a, st) : launch_trip<false, CfgB>(a, st); break;
    case 3: rc = rden ? launch_trip<true, CfgC>(a, st) : launch_trip<false, CfgC>(a, st); break;
    case 4: rc = rden ? launch_trip<true, CfgD>(a, st) : launch_trip<false, CfgD>(a, st); break;
    default: break;
  }
  if (rc != MBPLS_OK) return rc;
  MBPLS_RETURN_LAST();
}

int mbpls_fused_deflate_f64(double* Xt, long ld, int n, const double* ts, const double* rden_ts, const double* u0,
                            const double* u0u0, const double* rden_u0, const int* split_f0, const int* split_f1,
                            const int* split_block, int nsplit, int B, double* P_k, double* pss, double* w_next, double* norm_part,
                            double* Tnum, long ldt, void* stream) {
  if (!Xt || !ts || !split_f0 || !split_f1 || !split_block || !P_k || !pss || ld < n || B < 1) return MBPLS_ERR_ARG;
  if (u0 && (!u0u0 || !w_next || !norm_part || !Tnum || ldt < ld || (rden_ts && !rden_u0))) return MBPLS_ERR_ARG;
  if (nsplit == 0) return MBPLS_OK;
  FusedArgs a{nullptr, Xt, ld, n, u0, u0u0, ts, rden_ts, rden_u0, nullptr, nullptr, nullptr, split_f0, split_f1, split_block, nsplit, B,
              w_next, norm_part, Tnum, ldt, P_k, pss, nullptr};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = MBPLS_ERR_SIZE;
  switch (config_of(ld)) {
    case 1: rc = rden_ts ? launch_deflate<true, CfgA4>(a, st) : launch_deflate<false, CfgA4>(a, st); break;
    case 2: rc = rden_ts ? launch_deflate<true, CfgB>(a, st) : launch_deflate<false, CfgB>(a, st); break;
    case 3: rc = rden_ts ? launch_deflate<true, CfgC>(a, st) : launch_deflate<false, CfgC>(a, st); break;
    case 4: rc = rden_ts ? launch_deflate<true, CfgD>(a, st) : launch_deflate<false, CfgD>(a, st); break;
    default: break;
  }
  if (rc != MBPLS_OK) return rc;
  MBPLS_RETURN_LAST();
}

int mbpls_fused_deflate_rec_f64(double* Xt, long ld, int n, const double* ts, const double* rden_ts, const double* u0u0,
                                const double* rden_u0, const double* tsu0, const double* tsu0_masked, double* gdef,
                                const int* split_f0, const int* split_f1, const int* split_block, int nsplit, int B, double* P_k,
                                double* pss, double* w_next, double* norm_part, double* Tnum, long ldt, void* stream) {
  if (!Xt || !ts || !split_f0 || !split_f1 || !split_block || !P_k || !pss || ld < n || B < 1) return MBPLS_ERR_ARG;
  if (gdef && (!u0u0 || !tsu0 || !w_next || !norm_part || !Tnum || ldt < ld || (rden_ts && (!rden_u0 || !tsu0_masked))))
    return MBPLS_ERR_ARG;
  if (nsplit == 0) return MBPLS_OK;
  FusedArgs a{nullptr, Xt, ld, n, nullptr, u0u0, ts, rden_ts, rden_u0, tsu0_masked, tsu0, gdef, split_f0, split_f1, split_block, nsplit, B,
              w_next, norm_part, Tnum, ldt, P_k, pss, nullptr};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = MBPLS_ERR_SIZE;
  switch (config_of(ld)) {
    case 1: rc = rden_ts ? launch_deflate3<true, CfgA>(a, st) : launch_deflate3<false, CfgA>(a, st); break;
    case 2: rc = rden_ts ? launch_deflate3<true, CfgB>(a, st) : launch_deflate3<false, CfgB>(a, st); break;
    case 3: rc = rden_ts ? launch_deflate3<true, CfgC>(a, st) : launch_deflate3<false, CfgC>(a, st); break;
    case 4: rc = rden_ts ? launch_deflate3<true, CfgD>(a, st) : launch_deflate3<false, CfgD>(a, st); break;
    default: break;
  }
  if (rc != MBPLS_OK) return rc;
  MBPLS_RETURN_LAST();
}

}  // extern "C"
