// Probe: which cluster launch configurations does this box accept? (bring-up of tc_pair.cuh)
#include <cuda_runtime.h>
#include <stdio.h>
__global__ void __launch_bounds__(544, 1) k(int* out) {
  extern __shared__ char sm[];
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  if (threadIdx.x == 0) out[blockIdx.z * gridDim.y * gridDim.x + blockIdx.y * gridDim.x + blockIdx.x] = (int)r + (sm[0] & 0);
}
static void tryit(dim3 grid, int threads, size_t smem, dim3 cl) {
  int* d;
  cudaMalloc(&d, 4096);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 230400);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = dim3(threads, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cl.x; at[0].val.clusterDim.y = cl.y; at[0].val.clusterDim.z = cl.z;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, k, d);
  cudaError_t e2 = cudaDeviceSynchronize();
  int h[16] = {0};
  cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
  printf("grid (%u,%u,%u) threads %d smem %zu cluster (%u,%u,%u): launch %s, sync %s, ranks %d %d %d %d\n", grid.x, grid.y, grid.z,
         threads, smem, cl.x, cl.y, cl.z, cudaGetErrorString(e), cudaGetErrorString(e2), h[0], h[1], h[2], h[3]);
  cudaGetLastError();
  cudaFree(d);
}
int main() {
  tryit(dim3(1, 8, 2), 544, 189440, dim3(1, 2, 1));
  tryit(dim3(1, 8, 2), 544, 1024, dim3(1, 2, 1));
  tryit(dim3(1, 8, 2), 512, 189440, dim3(1, 2, 1));
  tryit(dim3(2, 4, 2), 544, 189440, dim3(2, 1, 1));
  tryit(dim3(8, 1, 2), 544, 189440, dim3(2, 1, 1));
  tryit(dim3(8, 1, 2), 256, 1024, dim3(2, 1, 1));
  return 0;
}
