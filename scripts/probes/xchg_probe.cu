// Where do the ~80-110 us of one xchg_epilogue_kernel launch go?  Single GPU (world = 1), synthetic partials, phase timestamps
// from %globaltimer (csrc/nipals.cu XSTAMP, compiled in here with -DMBPLS_XCHG_STAMPS).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DMBPLS_XCHG_STAMPS -o scripts/probes/xchg_probe scripts/probes/xchg_probe.cu
//   scripts/probes/xchg_probe <n> <B> <q> <nsplit> [ctas]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../mbpls_b200/csrc/nipals.cu"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__);          \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

template <class T>
T* dalloc(size_t n, double fill, bool randomize) {
  T* d;
  cudaMalloc(&d, n * sizeof(T));
  std::vector<T> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = randomize ? static_cast<T>(fill * (0.5 + (rand() % 1000) / 1000.0)) : static_cast<T>(fill);
  cudaMemcpy(d, h.data(), n * sizeof(T), cudaMemcpyHostToDevice);
  return d;
}

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 10000, B = argc > 2 ? atoi(argv[2]) : 4, q = argc > 3 ? atoi(argv[3]) : 1;
  const int nsplit = argc > 4 ? atoi(argv[4]) : 148, ctas = argc > 5 ? atoi(argv[5]) : 0;
  const long ld = (n + 15) / 16 * 16;
  std::vector<int> bso(B + 1);
  for (int b = 0; b <= B; ++b) bso[b] = static_cast<int>(static_cast<long>(nsplit) * b * (b + 1) / (static_cast<long>(B) * (B + 1)));  // blocks 1:2:..:B
  int* d_bso;
  cudaMalloc(&d_bso, (B + 1) * sizeof(int));
  cudaMemcpy(d_bso, bso.data(), (B + 1) * sizeof(int), cudaMemcpyHostToDevice);
  mbpls_xchg_args x = {};
  mbpls_epilogue_args& a = x.epi;
  a.n = n; a.B = B; a.q = q; a.nanmode = 0; a.norm_kind = 0; a.ldt = ld; a.ldf = ld; a.max_tol = -1.0;  // never converges: every launch does the full work
  a.red = dalloc<double>(static_cast<size_t>(B) * ld + B, 0.0, false);
  a.Yt = dalloc<double>(static_cast<size_t>(q) * ld, 1.0, true);
  a.T = dalloc<double>(static_cast<size_t>(B) * ld, 0.0, false);
  a.u = dalloc<double>(ld, 1.0, true);
  a.ts = dalloc<double>(ld, 0.0, false);
  a.ts_old = dalloc<double>(ld, 0.0, false);
  a.a = dalloc<double>(B, 0.0, false);
  a.v = dalloc<double>(q, 0.0, false);
  a.scal = dalloc<double>(MBPLS_SCAL_COUNT, 1.0, false);
  a.ctrl = dalloc<int>(MBPLS_CTRL_COUNT, 0.0, false);
  x.Tnum = dalloc<double>(static_cast<size_t>(nsplit) * ld, 1.0, true);
  x.ldp = ld;
  x.block_split_off = d_bso;
  x.norm_part = dalloc<double>(static_cast<size_t>(nsplit) * B, 1.0, true);
  x.n_norm_parts = nsplit;
  x.world = 1; x.rank = 0;
  x.counters = dalloc<unsigned int>(2, 0.0, false);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int it = 0; it < 5; ++it) CK(mbpls_nipals_xchg_epilogue_f64(&x, ctas, nullptr) ? cudaErrorUnknown : cudaSuccess);
  CK(cudaDeviceSynchronize());
  const int reps = 50;
  cudaEventRecord(e0);
  for (int it = 0; it < reps; ++it) mbpls_nipals_xchg_epilogue_f64(&x, ctas, nullptr);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long st[16] = {}, zero[16] = {};
  double acc[16] = {};
  for (int it = 0; it < 10; ++it) {  // phase stamps of isolated launches
    cudaMemcpyToSymbol(g_xchg_stamps, zero, sizeof(zero));
    mbpls_nipals_xchg_epilogue_f64(&x, ctas, nullptr);
    CK(cudaDeviceSynchronize());
    cudaMemcpyFromSymbol(st, g_xchg_stamps, sizeof(st));
    const int ks[] = {1, 3, 4, 5, 6, 7, 9};
    unsigned long long prev = st[0];
    for (int k : ks) {
      acc[k] += (st[k] - prev) / 10.0;
      prev = st[k];
    }
  }
  printf("n=%d B=%d q=%d nsplit=%d ctas=%d : %.1f us per back-to-back launch; phases (us): A split sums %.1f | arrive %.1f | last-CTA hand-off %.1f | "
         "epi 1 block scores+superweights %.1f | epi 2-3 ts, diff %.1f | epi 4 v %.1f | epi 5 u %.1f\n",
         n, B, q, nsplit, ctas, ms * 1e3 / reps, acc[1] / 1e3, acc[3] / 1e3, acc[4] / 1e3, acc[5] / 1e3, acc[6] / 1e3, acc[7] / 1e3, acc[9] / 1e3);
  return 0;
}
