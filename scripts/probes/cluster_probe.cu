// How many thread-block clusters of size 2 / 4 / 8 with ~200 KB of shared memory per CTA can be resident on this GPU?
//   nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/probes/cluster_probe scripts/probes/cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) dummy(double* p) {
  extern __shared__ double s[];
  s[threadIdx.x] = 1.0;
  __syncthreads();
  if (p) p[blockIdx.x] = s[0];
}
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs = 1; cs <= 16; cs *= 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms / cs * cs);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster size %2d: max active clusters %d (%d CTAs of %d SMs) %s\n", cs, n, n * cs, sms, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
