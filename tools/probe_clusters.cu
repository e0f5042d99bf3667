// Probe (debugging aid): how many thread-block clusters of size 1/2/4/8 can be co-resident on this GPU for a kernel that takes a
// whole SM's shared memory (the wgrad configuration)?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o probe_clusters probe_clusters.cu
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void k(int *out) {
  extern __shared__ char s[];
  if (threadIdx.x == 0 && out) out[blockIdx.x] = s[0];
}
int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("%s: %d SMs\n", p.name, p.multiProcessorCount);
  for (int smem : {220 * 1024, 100 * 1024, 16 * 1024}) {
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int cs : {1, 2, 4, 8}) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(148 / cs * cs);
      cfg.blockDim = dim3(192);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = -1;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
      printf("smem %3d KB cluster %d: max active clusters %d (%d CTAs)  %s\n", smem / 1024, cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
  }
  return 0;
}
