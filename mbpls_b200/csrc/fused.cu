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
//
// Cluster variants (template parameter CL): a feature is split by SAMPLES over the two CTAs of a thread-block cluster.
// Each CTA keeps only its half of the resident n-vectors (ts, u0 / u), which frees shared memory for the ring -- at
// n = 10,000 the deflation kernel's ring grows from 64 KB (16 KB chunks) to 120 KB (20 KB chunks x 3 x 2 workers) -- and
// lifts the longest one-pass feature from 10,240 to 20,480 samples.  The two halves of a feature meet in its dot products:
// after the worker-level reduction, thread 0 of the worker sends its partial into the peer CTA's shared memory with
// st.async (data + mbarrier complete_tx in one message over DSMEM), every thread waits on the local mbarrier and adds
// own + peer (commutative, so both CTAs hold bit-identical sums).  Two slots per worker alternate; a slot is rewritten only
// after the peer has passed the next exchange, which in turn needs this worker's next send, which follows the named barrier
// that every thread reaches after reading the slot -- no further handshake.  Outputs per feature (w~, p, p^2) are written by
// cluster rank 0 only; score partials by the CTA that owns the samples.
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
  const int* done;  // trip: skip the launch once *done is set; deflate: (may be null) run only if *done is set
  int sync_mode;  // stage hand-off protocol, see arrive_stage(): 2 = "empty" mbarrier (default), 0 / 1 = shared-memory counters
};

// TG threads per worker, EPTC units per thread per chunk, at most CPF chunks per feature, S ring stages
template <int TG, int EPTC_, int CPF_, int S_>
struct Cfg {
  static constexpr int kTG = TG, EPTC = EPTC_, CPF = CPF_, S = S_;
  static constexpr int G = 512 / TG, EPT = EPTC_ * CPF_, NW = TG / 32;
  static constexpr int UC = TG * EPTC_;  // units per (full) chunk
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
__device__ __forceinline__ unsigned atom_inc_smem_acqrel(unsigned* p) {
  unsigned old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(p)) : "memory");
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
  uint64_t* xbar;   // [G][2] mbarriers of the cluster exchange: the peer's partial has landed in xval
  double* xval;     // [G][2]
  __device__ Smem(unsigned char* base, long vlen, int nvec) {  // nvec vectors of vlen doubles in front of the ring
    vec0 = reinterpret_cast<double*>(base);
    vec1 = vec0 + vlen;
    ring = vec0 + static_cast<size_t>(nvec) * vlen;
    full = reinterpret_cast<uint64_t*>(ring + static_cast<size_t>(C::G) * C::S * 2 * C::UC);
    cnt = reinterpret_cast<unsigned*>(full + C::G * C::S);
    scratch = reinterpret_cast<double*>(full + 2 * C::G * C::S);
    xbar = reinterpret_cast<uint64_t*>(scratch + static_cast<size_t>(C::G) * 2 * 3 * C::NW);
    xval = reinterpret_cast<double*>(xbar + 2 * C::G);
  }
  __device__ __forceinline__ double* stage(int g, int s) const { return ring + (static_cast<size_t>(g) * C::S + s) * 2 * C::UC; }
};

template <class C>
size_t fused_smem_bytes(long vlen, int nvec) {
  return static_cast<size_t>(nvec) * vlen * 8 + static_cast<size_t>(C::G) * C::S * C::UC * 16 + static_cast<size_t>(C::G) * C::S * 16 +
         static_cast<size_t>(C::G) * 2 * 3 * C::NW * 8 + static_cast<size_t>(C::G) * 32 + 64;
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
__device__ __forceinline__ void init_sync(const Smem<C>& sm, int sync_mode) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < C::G * C::S; ++i) {
      mbar_init(&sm.full[i], 1);
      if (sync_mode == 2) mbar_init(reinterpret_cast<uint64_t*>(&sm.cnt[2 * i]), C::NW);  // "empty": one arrival per warp
      else sm.cnt[2 * i] = 0;
    }
    for (int i = 0; i < 2 * C::G; ++i) mbar_init(&sm.xbar[i], 1);
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

// The calling warp is done reading stage (g, s).  Returns (lane 0 only) whether it was the LAST warp of the worker to say so.
// Three hand-off protocols (FusedArgs::sync_mode, MBPLS_FUSED_SYNC):
//  0  relaxed shared-memory counter.  Every shared-memory read of the stage feeds a dot-product FMA that precedes this call in
//     program order, and an instruction cannot issue before its operands have arrived, so the reads have completed when the
//     counter moves -- an argument about the hardware, not the PTX memory model (racecheck reports the refill as a WAR hazard);
//  1  acq_rel counter + generic->async proxy fence in front of the refill: ordered by the memory model, invisible to racecheck;
//  2  an "empty" mbarrier per stage with one pending arrival per warp: mbarrier.arrive (release) returns the state before the
//     arrival, pending_count == 1 identifies the last arriver, whose test_wait (acquire) on the completed phase orders every
//     warp's reads before the bulk copy it then issues -- the consumer-release / producer-acquire of a TMA pipeline, with the
//     producer role falling to whichever warp arrives last.
template <class C>
__device__ __forceinline__ bool arrive_stage(int g, int s, int lane, const Smem<C>& sm, int sync_mode, uint32_t parity) {
  __syncwarp();
  if (lane != 0) return false;
  unsigned* c = &sm.cnt[2 * (g * C::S + s)];
  if (sync_mode == 2) {
    uint64_t* eb = reinterpret_cast<uint64_t*>(c);
    if (mbar_pending_count(mbar_arrive_state(eb)) != 1u) return false;
    mbar_wait(eb, parity);  // already complete: returns at once, with acquire semantics
    fence_proxy_async_smem();  // the reads were generic-proxy accesses, the refill is an async-proxy write
    return true;
  }
  const unsigned old = sync_mode ? atom_inc_smem_acqrel(c) : atom_inc_smem(c);
  if (old != C::NW - 1) return false;
  *reinterpret_cast<volatile unsigned*>(c) = 0;
  if (sync_mode) fence_proxy_async_smem();
  return true;
}

// The last warp of the worker to arrive on a stage (which held chunk c of feature j) refills it with the chunk S
// positions further down the worker's stream, immediately (a free stage is lost ring depth).
template <class C>
__device__ __forceinline__ void refill_if_last(bool last, const double* __restrict__ X, long ld, int units, int ncf, int g, int s,
                                               int j, int c, int f1, const Smem<C>& sm) {
  if (last) {
    int c2 = c + C::S, f2 = j;
    while (c2 >= ncf) { c2 -= ncf; ++f2; }
    if (f2 < f1) issue_chunk<C>(X, ld, units, g, s, f2, c2, sm);
  }
}

// Sum of the two CTAs' partials of one worker pair (cluster variants).  `xs` (slot) and `xp` (phase bits) are per-thread state.
template <class C>
__device__ __forceinline__ double pair_sum(double own, int g, int tg, uint32_t peer, int& xs, uint32_t& xp, const Smem<C>& sm) {
  uint64_t* bar = &sm.xbar[2 * g + xs];
  double* slot = &sm.xval[2 * g + xs];
  if (tg == 0) {
    mbar_arrive_expect_tx(bar, 8);
    st_async_f64(mapa_shared(smem_u32(slot), peer), own, mapa_shared(smem_u32(bar), peer));
  }
  // CTA-scope acquire: the value arrives through the mbarrier's transaction count (like a multicast bulk copy), and a
  // cluster-scope acquire would invalidate L1 on every exchange (CCTL.IVALL: 8 % of the stall samples in ncu)
  mbar_wait(bar, (xp >> xs) & 1u);
  const double other = *reinterpret_cast<volatile double*>(slot);
  xp ^= 1u << xs;
  xs ^= 1;
  return own + other;
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

// Geometry of one CTA's share of a feature: the whole feature, or (cluster variants) the half owned by this cluster rank.
struct Geo {
  int units;      // 16-byte units of a feature handled by this CTA
  int uoff;       // first unit (0, or units_total / 2 for cluster rank 1)
  int nloc;       // samples of this CTA's share that are real (< n)
  int wbase;      // first worker (split) index of this CTA
  uint32_t peer;  // cluster rank of the other CTA
  bool lead;      // this CTA writes the per-feature outputs
};
template <class C, bool CL>
__device__ __forceinline__ Geo make_geo(long ld, int n) {
  Geo ge;
  const int total = static_cast<int>(ld >> 1);
  if (CL) {
    const uint32_t r = cluster_ctarank();
    ge.units = total >> 1;  // ld % 16 == 0, so the halves are equal and 64-byte aligned
    ge.uoff = static_cast<int>(r) * ge.units;
    ge.wbase = static_cast<int>(blockIdx.x >> 1) * C::G;
    ge.peer = r ^ 1u;
    ge.lead = r == 0;
  } else {
    ge.units = total;
    ge.uoff = 0;
    ge.wbase = static_cast<int>(blockIdx.x) * C::G;
    ge.peer = 0;
    ge.lead = true;
  }
  ge.nloc = max(0, min(2 * ge.units, n - 2 * ge.uoff));
  return ge;
}

template <bool NANMODE, class C, bool CL>
__device__ __forceinline__ void fused_trip_body(const FusedArgs& a, const Smem<C>& sm, const Geo& ge) {
  const long ld = a.ld;
  const int units = ge.units;
  const int ncf = (units + C::UC - 1) / C::UC;
  const int g = threadIdx.x / C::kTG, tg = threadIdx.x % C::kTG;
  const int lane = threadIdx.x & 31, wig = tg >> 5;
  const int wk = ge.wbase + g;
  if (wk >= a.nsplit) return;
  const double* __restrict__ X = a.Xt + 2 * static_cast<size_t>(ge.uoff);  // this CTA's share of every feature
  const int f0 = a.split_f0[wk], f1 = a.split_f1[wk];
  if (tg == 0) prime_ring<C>(X, ld, units, ncf, g, f0, f1, sm);
  const double inv_uu = 1.0 / *a.uu;
  const double2* __restrict__ u2 = reinterpret_cast<const double2*>(sm.vec0);
  double* scratch = sm.scratch + static_cast<size_t>(g) * 2 * 3 * C::NW;
  const int sync_mode = a.sync_mode;

  double2 acc[C::EPT], x[C::EPT];
#pragma unroll
  for (int k = 0; k < C::EPT; ++k) acc[k] = x[k] = make_double2(0.0, 0.0);
  double normsq = 0.0, wj = 0.0;
  int s = 0;
  uint32_t ph = 0;
  int flip = 0;
  int xs = 0;
  uint32_t xp = 0;

  for (int j = f0; j <= f1; ++j) {
    const bool load = j < f1;  // the last iteration only applies the last weight
    const double rd = (NANMODE && load) ? a.rden[j] : inv_uu;  // 1 / (masked) u'u of this feature
    double numa = 0.0, numb = 0.0, numc = 0.0, numd = 0.0;
    int s_use = s;  // s: next stage to load from; s_use: stage of the chunk being consumed
    uint32_t ph_use = ph;

    // previous feature's weight into the accumulators, then this feature's chunk c into the freed registers
    auto load_chunk = [&](const int c) {
      const bool have = load && c < ncf;
      const double2* __restrict__ xs_ = reinterpret_cast<const double2*>(sm.stage(g, s));
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
          xv = xs_[l];
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
      refill_if_last<C>(arrive_stage<C>(g, s_use, lane, sm, sync_mode, ph_use), X, ld, units, ncf, g, s_use, j, c, f1, sm);
      if (++s_use == C::S) { s_use = 0; ph_use ^= 1u; }
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
    if (CL) one[0] = pair_sum<C>(one[0], g, tg, ge.peer, xs, xp, sm);
    wj = one[0] * rd;
    if (tg == 0 && ge.lead) a.w[j] = wj;
    normsq = fma(wj, wj, normsq);
  }

  // partial block scores of this split (the layout xw_kernel writes)
  double2* tn = reinterpret_cast<double2*>(a.Tnum + static_cast<size_t>(wk) * a.ldt) + ge.uoff;
#pragma unroll
  for (int k = 0; k < C::EPT; ++k) {
    const int gi = (k / C::EPTC) * C::UC + tg + (k % C::EPTC) * C::kTG;
    if (gi < units) tn[gi] = acc[k];
  }
  if (tg == 0 && ge.lead) a.norm_part[static_cast<size_t>(wk) * a.B + a.split_block[wk]] = normsq;
}

template <bool NANMODE, class C, bool CL>
__global__ void __launch_bounds__(512, 1) fused_trip_kernel(const FusedArgs a) {
  if (a.done && *a.done) return;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Geo ge = make_geo<C, CL>(a.ld, a.n);
  const Smem<C> sm(smem_raw, 2 * ge.units, 1);  // u (this CTA's share)
  init_sync<C>(sm, a.sync_mode);
  for (int i = threadIdx.x; i < 2 * ge.units; i += blockDim.x) sm.vec0[i] = i < ge.nloc ? a.u[2 * ge.uoff + i] : 0.0;
  __syncthreads();
  if (CL) cluster_sync_all();  // the peer's exchange barriers are initialised before anything is sent to them
  fused_trip_body<NANMODE, C, CL>(a, sm, ge);
  if (CL) cluster_sync_all();  // no CTA exits while its peer may still address its shared memory
}

// ------------------------------------------------------------------------------------------
// loadings + deflation + the whole first trip of the next component; same pipeline: the score update of
// feature j-1 rides on the load phase of feature j.  NaN mode: NaN entries take part as zeros and are written
// back as NaN (:969 keeps them); loadings and next weights use the masked reciprocal denominators rden / rden2.
// ------------------------------------------------------------------------------------------
template <bool NANMODE, class C, bool CL>
__device__ __forceinline__ void fused_deflate_body(const FusedArgs& a, const Smem<C>& sm, const Geo& ge) {
  const long ld = a.ld;
  const int units = ge.units;
  const int ncf = (units + C::UC - 1) / C::UC;
  const bool next = a.u != nullptr;
  const int g = threadIdx.x / C::kTG, tg = threadIdx.x % C::kTG;
  const int lane = threadIdx.x & 31, wig = tg >> 5;
  const int wk = ge.wbase + g;
  if (wk >= a.nsplit) return;
  double* __restrict__ X = a.Xw + 2 * static_cast<size_t>(ge.uoff);  // this CTA's share of every feature
  const int f0 = a.split_f0[wk], f1 = a.split_f1[wk];
  if (tg == 0) prime_ring<C>(X, ld, units, ncf, g, f0, f1, sm);
  const double inv_uu = next ? 1.0 / *a.uu : 1.0;
  const double2* __restrict__ ts2 = reinterpret_cast<const double2*>(sm.vec0);
  const double2* __restrict__ u2 = reinterpret_cast<const double2*>(sm.vec1);
  double* scratch = sm.scratch + static_cast<size_t>(g) * 2 * 3 * C::NW;
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  const int sync_mode = a.sync_mode;

  double2 acc[C::EPT], x[C::EPT];
#pragma unroll
  for (int k = 0; k < C::EPT; ++k) acc[k] = x[k] = make_double2(0.0, 0.0);
  double normsq = 0.0, wj = 0.0;
  int s = 0;
  uint32_t ph = 0;
  int flip = 0;
  int xs = 0;
  uint32_t xp = 0;

  for (int j = f0; j <= f1; ++j) {
    const bool load = j < f1;
    const double rdp = (NANMODE && load) ? a.rden[j] : 1.0;                      // loadings: 1 / masked ts'ts (dense: not divided, :920)
    const double rdw = (NANMODE && load && next) ? a.rden2[j] : inv_uu;          // next weights: 1 / masked u0'u0
    double pa = 0.0, pb = 0.0, pc = 0.0, pd = 0.0;
    uint32_t mx = 0, my = 0;  // NaN mode: which of this thread's entries are NaN (they must be stored back as NaN)
    int s_use = s;
    uint32_t ph_use = ph;

    auto load_chunk = [&](const int c) {
      const bool have = load && c < ncf;
      const double2* __restrict__ xs_ = reinterpret_cast<const double2*>(sm.stage(g, s));
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
          xv = xs_[l];
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
      refill_if_last<C>(arrive_stage<C>(g, s_use, lane, sm, sync_mode, ph_use), X, ld, units, ncf, g, s_use, j, c, f1, sm);
      if (++s_use == C::S) { s_use = 0; ph_use ^= 1u; }
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
    if (CL) one[0] = pair_sum<C>(one[0], g, tg, ge.peer, xs, xp, sm);
    const double pj = one[0] * rdp;
    double2* __restrict__ xg = reinterpret_cast<double2*>(X + static_cast<size_t>(j) * ld);
    double wa = 0.0, wb = 0.0, wc = 0.0, wd = 0.0;
#pragma unroll
    for (int k = 0; k < C::EPT; ++k) {
      const int gi = (k / C::EPTC) * C::UC + tg + (k % C::EPTC) * C::kTG;
      if (((k / C::EPTC) + 1 < ncf) || gi < units) {
        const double2 tv = ts2[gi];
        const double2 uv = u2[gi];  // issued together with the ts load (zeros when there is no next component)
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
      if (CL) two[0] = pair_sum<C>(two[0], g, tg, ge.peer, xs, xp, sm);
      wj = two[0] * rdw;
      normsq = fma(wj, wj, normsq);
    }
    if (tg == 0 && ge.lead) {
      a.P_k[j] = pj;
      a.pss[j] = pj * pj;
      if (next) a.w[j] = wj;
    }
  }
  if (!next) return;
  double2* tn = reinterpret_cast<double2*>(a.Tnum + static_cast<size_t>(wk) * a.ldt) + ge.uoff;
#pragma unroll
  for (int k = 0; k < C::EPT; ++k) {
    const int gi = (k / C::EPTC) * C::UC + tg + (k % C::EPTC) * C::kTG;
    if (gi < units) tn[gi] = acc[k];
  }
  if (tg == 0 && ge.lead) a.norm_part[static_cast<size_t>(wk) * a.B + a.split_block[wk]] = normsq;
}

template <bool NANMODE, class C, bool CL>
__global__ void __launch_bounds__(512, 1) fused_deflate_kernel(const FusedArgs a) {
  if (a.done && !*a.done) return;  // enqueued behind trips that have not converged yet: nothing to close (see engine.nipals_fit)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Geo ge = make_geo<C, CL>(a.ld, a.n);
  const Smem<C> sm(smem_raw, 2 * ge.units, 2);  // ts | u0 (this CTA's share)
  init_sync<C>(sm, a.sync_mode);
  const bool next = a.u != nullptr;
  for (int i = threadIdx.x; i < 2 * ge.units; i += blockDim.x) {
    sm.vec0[i] = i < ge.nloc ? a.ts[2 * ge.uoff + i] : 0.0;
    sm.vec1[i] = (next && i < ge.nloc) ? a.u[2 * ge.uoff + i] : 0.0;
  }
  __syncthreads();
  if (CL) cluster_sync_all();
  fused_deflate_body<NANMODE, C, CL>(a, sm, ge);
  if (CL) cluster_sync_all();
}

// ------------------------------------------------------------------------------------------
// StandardScaler.fit_transform (mbpls.py:307,314) AND the first trip of the first component in one read + one write.
// The first trip's u is the first (standardised) Y column -- known before X is touched -- so while a feature sits in
// registers for its statistics it can also deliver its first weight w~_j = z_j . u0 / u0'u0 and add to the first block-score
// partials, exactly as the deflation pass does for later components.  Two worker reductions per feature: the sum (mean), then
// {sum of deviations, sum of squared deviations, deviations . u0} -- the corrected two-pass variance of sklearn's
// _incremental_mean_and_var, and, because z = d / scale, also sum z^2 and z . u0 without a third one.  Dense data only (a NaN shows up as a non-finite mean and is reported by the caller, like check_array).
// Samples beyond n (the zero padding of a feature) take no part and stay zero.
// ------------------------------------------------------------------------------------------
struct StdArgs {
  double* mean;
  double* var;
  double* scale;
  long long* seen;
  double* zss;
};

template <class C>
__global__ void __launch_bounds__(512, 1) fused_standardize_kernel(const FusedArgs a, const StdArgs so) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const Geo ge = make_geo<C, false>(a.ld, a.n);
  const Smem<C> sm(smem_raw, 2 * ge.units, 1);  // u0
  init_sync<C>(sm, a.sync_mode);
  for (int i = threadIdx.x; i < 2 * ge.units; i += blockDim.x) sm.vec0[i] = i < a.n ? a.u[i] : 0.0;
  __syncthreads();
  const long ld = a.ld;
  const int units = ge.units;
  const int ncf = (units + C::UC - 1) / C::UC;
  const int g = threadIdx.x / C::kTG, tg = threadIdx.x % C::kTG;
  const int lane = threadIdx.x & 31, wig = tg >> 5;
  const int wk = ge.wbase + g;
  if (wk >= a.nsplit) return;
  double* __restrict__ X = a.Xw;
  const int f0 = a.split_f0[wk], f1 = a.split_f1[wk];
  if (tg == 0) prime_ring<C>(X, ld, units, ncf, g, f0, f1, sm);
  const double inv_uu = 1.0 / *a.uu;
  const double cnt = static_cast<double>(a.n);
  const double2* __restrict__ u2 = reinterpret_cast<const double2*>(sm.vec0);
  double* scratch = sm.scratch + static_cast<size_t>(g) * 2 * 3 * C::NW;
  const int sync_mode = a.sync_mode;
  const int nn = a.n;  // register slots at or beyond sample n (padding of the feature, or beyond it) hold no data
  double2 acc[C::EPT], x[C::EPT];
#pragma unroll
  for (int k = 0; k < C::EPT; ++k) acc[k] = x[k] = make_double2(0.0, 0.0);
  double normsq = 0.0, wj = 0.0;
  int s = 0;
  uint32_t ph = 0;
  int flip = 0;

  for (int j = f0; j <= f1; ++j) {
    const bool load = j < f1;
    double sa = 0.0, sb = 0.0, sc = 0.0, sd = 0.0;
    int s_use = s;
    uint32_t ph_use = ph;
#pragma unroll
    for (int c = 0; c < C::CPF; ++c) {
      const bool have = load && c < ncf;
      const double2* __restrict__ xs_ = reinterpret_cast<const double2*>(sm.stage(g, s));
      if (have) {
        mbar_wait(&sm.full[g * C::S + s], ph);
        if (++s == C::S) { s = 0; ph ^= 1u; }
      }
      const bool whole = c + 1 < ncf;
#pragma unroll
      for (int e = 0; e < C::EPTC; ++e) {
        const int l = tg + e * C::kTG, gi = c * C::UC + l, k = c * C::EPTC + e;
        acc[k].x = fma(wj, x[k].x, acc[k].x);  // first block-score partials with the previous feature's weight
        acc[k].y = fma(wj, x[k].y, acc[k].y);
        double2 xv = make_double2(0.0, 0.0);
        if (have && (whole || gi < units)) xv = xs_[l];
        if (e & 1) {
          sc += xv.x;
          sd += xv.y;
        } else {
          sa += xv.x;
          sb += xv.y;
        }
        x[k] = xv;
      }
      if (have) {
        refill_if_last<C>(arrive_stage<C>(g, s_use, lane, sm, sync_mode, ph_use), X, ld, units, ncf, g, s_use, j, c, f1, sm);
        if (++s_use == C::S) { s_use = 0; ph_use ^= 1u; }
      }
    }
    if (!load) break;
    double one[1] = {(sa + sb) + (sc + sd)};
    worker_sum<1, C::kTG>(one, scratch + flip * 3 * C::NW, g, wig, lane);
    flip ^= 1;
    const double mean = one[0] / cnt;
    double thr[3] = {0.0, 0.0, 0.0};  // sum of deviations (correction term), sum of squared deviations, deviations . u0
#pragma unroll
    for (int k = 0; k < C::EPT; ++k) {
      const int gi = (k / C::EPTC) * C::UC + tg + (k % C::EPTC) * C::kTG;
      const int e0 = 2 * gi;
      double2 d;
      d.x = e0 < nn ? x[k].x - mean : 0.0;
      d.y = e0 + 1 < nn ? x[k].y - mean : 0.0;
      if (e0 < nn) {
        const double2 uv = u2[gi];
        thr[2] = fma(d.x, uv.x, thr[2]);
        thr[2] = fma(d.y, uv.y, thr[2]);
      }
      thr[0] += d.x + d.y;
      thr[1] = fma(d.x, d.x, thr[1]);
      thr[1] = fma(d.y, d.y, thr[1]);
      x[k] = d;
    }
    worker_sum<3, C::kTG>(thr, scratch + flip * 3 * C::NW, g, wig, lane);
    flip ^= 1;
    const double var = (thr[1] - thr[0] * thr[0] / cnt) / cnt;
    const double eps = 2.220446049250313e-16;
    const double bound = cnt * eps * var + (cnt * mean * eps) * (cnt * mean * eps);
    const double scale = (var <= bound) ? 1.0 : sqrt(var);  // _is_constant_feature -> scale 1
    // z = d / scale: sum z^2 = sum d^2 / scale^2 and z . u0 = (d . u0) / scale need no third reduction
    wj = thr[2] / scale * inv_uu;
    normsq = fma(wj, wj, normsq);
    double2* __restrict__ xg = reinterpret_cast<double2*>(X + static_cast<size_t>(j) * ld);
    const UniformDivisor by_scale(scale);  // (x - mean) / scale, bit-identical to StandardScaler's true division (common.cuh)
#pragma unroll
    for (int k = 0; k < C::EPT; ++k) {
      const int gi = (k / C::EPTC) * C::UC + tg + (k % C::EPTC) * C::kTG;
      if (((k / C::EPTC) + 1 < ncf) || gi < units) {
        double2 z;
        z.x = by_scale(x[k].x);
        z.y = by_scale(x[k].y);
        st_stream(xg + gi, z);
        x[k] = z;
      }
    }
    if (tg == 0) {
      so.mean[j] = mean;
      so.var[j] = var;
      so.scale[j] = scale;
      so.seen[j] = static_cast<long long>(a.n);
      so.zss[j] = thr[1] / (scale * scale);
      a.w[j] = wj;
    }
  }
  double2* tn = reinterpret_cast<double2*>(a.Tnum + static_cast<size_t>(wk) * a.ldt);
#pragma unroll
  for (int k = 0; k < C::EPT; ++k) {
    const int gi = (k / C::EPTC) * C::UC + tg + (k % C::EPTC) * C::kTG;
    if (gi < units) tn[gi] = acc[k];
  }
  if (tg == 0) a.norm_part[static_cast<size_t>(wk) * a.B + a.split_block[wk]] = normsq;
}

// Configurations by feature length (units = 16-byte units per feature handled by ONE CTA <= TG*EPTC*CPF).  Measured
// (profiles/r1_notes.md): every chunk costs a worker ~0.2 us of handshakes (wait, counter, refill), so chunks are as large as
// the ring allows: the 16 KB x 8 ring ran the n = 10,000 trip at 4.9 TB/s, the 40 KB x 3 ring runs it at 6.6 TB/s.
using CfgA = Cfg<512, 5, 2, 3>;   // <= 5120 units: one worker per CTA, 40 KB chunks (u + ring = 200 KB)
using CfgA4 = Cfg<512, 2, 5, 4>;  // same length with two resident vectors (deflate: ts and u0): only 64 KB of ring left
using CfgB = Cfg<256, 5, 2, 3>;   // <= 2560 units: two workers, 20 KB chunks
using CfgC = Cfg<128, 5, 2, 4>;   // <= 1280 units: four workers, 10 KB chunks
using CfgD = Cfg<64, 10, 1, 2>;   // <= 640 units: eight workers, one chunk per feature

// MBPLS_FUSED_CLUSTER=0 keeps features of up to 10,240 samples on the single-CTA deflation kernel (A/B measurements);
// MBPLS_FUSED_SYNC=0 / 1 select the counter-based stage hand-offs instead of the default "empty" mbarrier (FusedArgs::sync_mode)
int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
bool cluster_enabled() {
  static int v = -1;
  if (v < 0) v = env_int("MBPLS_FUSED_CLUSTER", 1) != 0;
  return v != 0;
}
int default_sync_mode() {
  static int v = -1;
  if (v < 0) {
    v = env_int("MBPLS_FUSED_SYNC", 2);
    if (v < 0 || v > 2) v = 2;
  }
  return v;
}

template <class K>
int launch_fused(K kernel, const FusedArgs& a, int workers_per_cta, size_t smem, bool cluster, cudaStream_t st) {
  if (smem > static_cast<size_t>(smem_optin())) return MBPLS_ERR_SIZE;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  const int groups = (a.nsplit + workers_per_cta - 1) / workers_per_cta;  // CTAs, or CTA pairs
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(cluster ? 2 * groups : groups), 1, 1);
  cfg.blockDim = dim3(512, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cluster ? 1 : 0;
  const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, a);
  return e == cudaSuccess ? MBPLS_OK : MBPLS_CUDA_ERR(e);
}

// CTA pairs of a cluster kernel that can be resident at once (the one-pass kernels are persistent: one wave only)
template <class K>
int max_active_pairs(K kernel, size_t smem) {
  if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(num_sms() / 2 * 2), 1, 1);
  cfg.blockDim = dim3(512, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kernel, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

// Kernel plan for a leading dimension.  trip / deflate: configuration letter, cluster flag; workers: splits the split table
// should hold (= persistent workers of one wave).  A feature of up to 10,240 samples keeps the single-CTA trip kernel (its ring
// is already 120 KB) but deflates through CTA pairs (ts and u0 halved per CTA: 120 KB of ring instead of 64 KB); features of up
// to 20,480 samples run both kernels on CTA pairs.
struct Plan {
  int trip = 0, deflate = 0;  // 0: unsupported; 1..4 = CfgA..CfgD per CTA; 5 = CfgA4
  bool trip_cl = false, deflate_cl = false;
  int workers = 0;
};

Plan plan_of(long ld) {
  Plan pl;
  if (ld < 16 || (ld % 16) != 0) return pl;
  const long units = ld >> 1;
  const int sms = num_sms();
  if (units <= 640) { pl.trip = pl.deflate = 4; pl.workers = sms * CfgD::G; return pl; }
  if (units <= 1280) { pl.trip = pl.deflate = 3; pl.workers = sms * CfgC::G; return pl; }
  if (units <= 2560) { pl.trip = pl.deflate = 2; pl.workers = sms * CfgB::G; return pl; }
  static int pairs_b = -1, pairs_a = -1, pairs_a4 = -1;  // resident CTA pairs of the three cluster kernels (device 0's answer is reused)
  if (units <= 5120) {
    pl.trip = 1;
    pl.deflate = 5;
    pl.workers = sms * CfgA::G;
    if (cluster_enabled()) {
      if (pairs_b < 0) pairs_b = max_active_pairs(fused_deflate_kernel<false, CfgB, true>, fused_smem_bytes<CfgB>(ld >> 1, 2));
      if (pairs_b * CfgB::G >= pl.workers * 9 / 10) {  // (almost) every SM gets a CTA: worth it
        pl.deflate = 2;
        pl.deflate_cl = true;
        pl.workers = min(pl.workers, pairs_b * CfgB::G);
      }
    }
    return pl;
  }
  if (units <= 10240 && cluster_enabled()) {
    if (pairs_a < 0) pairs_a = max_active_pairs(fused_trip_kernel<false, CfgA, true>, fused_smem_bytes<CfgA>(ld >> 1, 1));
    if (pairs_a4 < 0) pairs_a4 = max_active_pairs(fused_deflate_kernel<false, CfgA4, true>, fused_smem_bytes<CfgA4>(ld >> 1, 2));
    const int pairs = min(pairs_a, pairs_a4);
    if (pairs > 0) {
      pl.trip = 1;
      pl.deflate = 5;
      pl.trip_cl = pl.deflate_cl = true;
      pl.workers = pairs * CfgA::G;
    }
  }
  return pl;
}

template <bool NANMODE, bool CL>
int dispatch_trip(int cfg, const FusedArgs& a, cudaStream_t st) {
  const long vlen = CL ? a.ld >> 1 : a.ld;
  switch (cfg) {
    case 1: return launch_fused(fused_trip_kernel<NANMODE, CfgA, CL>, a, CfgA::G, fused_smem_bytes<CfgA>(vlen, 1), CL, st);
    case 2: return launch_fused(fused_trip_kernel<NANMODE, CfgB, CL>, a, CfgB::G, fused_smem_bytes<CfgB>(vlen, 1), CL, st);
    case 3: return launch_fused(fused_trip_kernel<NANMODE, CfgC, false>, a, CfgC::G, fused_smem_bytes<CfgC>(a.ld, 1), false, st);
    case 4: return launch_fused(fused_trip_kernel<NANMODE, CfgD, false>, a, CfgD::G, fused_smem_bytes<CfgD>(a.ld, 1), false, st);
    default: return MBPLS_ERR_SIZE;
  }
}

template <bool NANMODE, bool CL>
int dispatch_deflate(int cfg, const FusedArgs& a, cudaStream_t st) {
  const long vlen = CL ? a.ld >> 1 : a.ld;
  switch (cfg) {
    case 5: return launch_fused(fused_deflate_kernel<NANMODE, CfgA4, CL>, a, CfgA4::G, fused_smem_bytes<CfgA4>(vlen, 2), CL, st);
    case 2: return launch_fused(fused_deflate_kernel<NANMODE, CfgB, CL>, a, CfgB::G, fused_smem_bytes<CfgB>(vlen, 2), CL, st);
    case 3: return launch_fused(fused_deflate_kernel<NANMODE, CfgC, false>, a, CfgC::G, fused_smem_bytes<CfgC>(a.ld, 2), false, st);
    case 4: return launch_fused(fused_deflate_kernel<NANMODE, CfgD, false>, a, CfgD::G, fused_smem_bytes<CfgD>(a.ld, 2), false, st);
    default: return MBPLS_ERR_SIZE;
  }
}

}  // namespace

extern "C" {

/* persistent workers of the one-pass kernels per PAIR of SMs for this leading dimension (pure function of ld, no device
 * needed): 16 / 8 / 4 / 2 for features of up to 1280 / 2560 / 5120 / 10240 samples, 1 up to 20480 (a feature is then split
 * over the CTA pair of a cluster), 0 = too long: use the two-pass kernels */
int mbpls_fused_workers_per_sm_pair(long ld) {
  if (ld < 16 || (ld % 16) != 0) return 0;
  const long units = ld >> 1;
  return units <= 640 ? 16 : units <= 1280 ? 8 : units <= 2560 ? 4 : units <= 5120 ? 2 : units <= 10240 ? 1 : 0;
}

/* splits (persistent workers of one wave) to size the split table to on the current device; 0 = feature too long */
int mbpls_fused_total_workers(long ld) { return plan_of(ld).workers; }

/* 1 if the deflation pass for this leading dimension runs on CTA pairs (thread-block clusters), else 0 */
int mbpls_fused_uses_clusters(long ld) {
  const Plan pl = plan_of(ld);
  return (pl.trip_cl ? 1 : 0) | (pl.deflate_cl ? 2 : 0);
}

int mbpls_nipals_fused_trip_f64(const double* Xt, long ld, int n, const double* u, const double* uu, const double* rden,
                                const int* split_f0, const int* split_f1, const int* split_block, int nsplit, int B, double* w,
                                double* norm_part, double* Tnum, long ldt, const int* done, void* stream) {
  if (!Xt || !u || !uu || !split_f0 || !split_f1 || !split_block || !w || !norm_part || !Tnum || ld < n || ldt < ld || B < 1)
    return MBPLS_ERR_ARG;
  if (nsplit == 0) return MBPLS_OK;
  const Plan pl = plan_of(ld);
  if (!pl.trip || nsplit > pl.workers) return MBPLS_ERR_SIZE;
  FusedArgs a{Xt, nullptr, ld, n, u, uu, nullptr, rden, nullptr, split_f0, split_f1, split_block, nsplit, B,
              w, norm_part, Tnum, ldt, nullptr, nullptr, done, default_sync_mode()};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pl.trip_cl) return rden ? dispatch_trip<true, true>(pl.trip, a, st) : dispatch_trip<false, true>(pl.trip, a, st);
  return rden ? dispatch_trip<true, false>(pl.trip, a, st) : dispatch_trip<false, false>(pl.trip, a, st);
}

int mbpls_fused_deflate_f64(double* Xt, long ld, int n, const double* ts, const double* rden_ts, const double* u0,
                            const double* u0u0, const double* rden_u0, const int* split_f0, const int* split_f1,
                            const int* split_block, int nsplit, int B, double* P_k, double* pss, double* w_next, double* norm_part,
                            double* Tnum, long ldt, const int* only_if_done, void* stream) {
  if (!Xt || !ts || !split_f0 || !split_f1 || !split_block || !P_k || !pss || ld < n || B < 1) return MBPLS_ERR_ARG;
  if (u0 && (!u0u0 || !w_next || !norm_part || !Tnum || ldt < ld || (rden_ts && !rden_u0))) return MBPLS_ERR_ARG;
  if (nsplit == 0) return MBPLS_OK;
  const Plan pl = plan_of(ld);
  if (!pl.deflate || nsplit > pl.workers) return MBPLS_ERR_SIZE;
  FusedArgs a{nullptr, Xt, ld, n, u0, u0u0, ts, rden_ts, rden_u0, split_f0, split_f1, split_block, nsplit, B,
              w_next, norm_part, Tnum, ldt, P_k, pss, only_if_done, default_sync_mode()};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (pl.deflate_cl) return rden_ts ? dispatch_deflate<true, true>(pl.deflate, a, st) : dispatch_deflate<false, true>(pl.deflate, a, st);
  return rden_ts ? dispatch_deflate<true, false>(pl.deflate, a, st) : dispatch_deflate<false, false>(pl.deflate, a, st);
}

/* StandardScaler.fit_transform of X in place (mbpls.py:307,314) fused with the complete first trip of the first component
 * (u = u0, the first standardised Y column): statistics out as for mbpls_standardize_fit_f64, plus w, norm_part, Tnum as for
 * mbpls_nipals_fused_trip_f64.  Dense data, features of up to 10,240 samples (returns MBPLS_ERR_SIZE otherwise). */
int mbpls_fused_standardize_f64(double* Xt, long ld, int n, const double* u0, const double* u0u0, const int* split_f0,
                                const int* split_f1, const int* split_block, int nsplit, int B, double* mean, double* var,
                                double* scale, long long* seen, double* zss, double* w, double* norm_part, double* Tnum, long ldt,
                                void* stream) {
  if (!Xt || !u0 || !u0u0 || !split_f0 || !split_f1 || !split_block || !mean || !var || !scale || !seen || !zss || !w ||
      !norm_part || !Tnum || ld < n || ldt < ld || B < 1)
    return MBPLS_ERR_ARG;
  if (nsplit == 0) return MBPLS_OK;
  const Plan pl = plan_of(ld);
  if (!pl.trip || pl.trip_cl || nsplit > pl.workers) return MBPLS_ERR_SIZE;
  FusedArgs a{nullptr, Xt, ld, n, u0, u0u0, nullptr, nullptr, nullptr, split_f0, split_f1, split_block, nsplit, B,
              w, norm_part, Tnum, ldt, nullptr, nullptr, nullptr, default_sync_mode()};
  const StdArgs so{mean, var, scale, seen, zss};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define STD_LAUNCH(CFG)                                                                                                      \
  do {                                                                                                                       \
    const size_t smem = fused_smem_bytes<CFG>(ld, 1);                                                                        \
    if (smem > static_cast<size_t>(smem_optin())) return MBPLS_ERR_SIZE;                                                     \
    cudaFuncSetAttribute(fused_standardize_kernel<CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)); \
    fused_standardize_kernel<CFG><<<(nsplit + CFG::G - 1) / CFG::G, 512, smem, st>>>(a, so);                                  \
  } while (0)
  switch (pl.trip) {
    case 1: STD_LAUNCH(CfgA); break;
    case 2: STD_LAUNCH(CfgB); break;
    case 3: STD_LAUNCH(CfgC); break;
    case 4: STD_LAUNCH(CfgD); break;
    default: return MBPLS_ERR_SIZE;
  }
#undef STD_LAUNCH
  MBPLS_RETURN_LAST();
}

}  // extern "C"
