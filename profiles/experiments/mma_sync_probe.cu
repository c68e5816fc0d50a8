// mma_sync_probe.cu -- throughput of the LEGACY warp-level tensor path on sm_100a (mma.sync.m16n8k8 tf32, fp32 accumulate)
// and of plain FFMA, registers only: what a one-CTA-per-client fused kernel could get without tcgen05 descriptors / TMEM.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mma_sync_probe mma_sync_probe.cu && ./mma_sync_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int NACC>
__global__ void mma_kernel(float* out, int iters) {
  float c[NACC][4];
  unsigned a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(1.0f + threadIdx.x * 1e-3f + i);
  for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(0.5f + threadIdx.x * 1e-3f + i);
  for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < NACC; ++j) mma_tf32(c[j], a, b);
  }
  float s = 0.f;
  for (int j = 0; j < NACC; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ffma_kernel(float* out, int iters) {
  float c[32];
  const float a = 1.0f + threadIdx.x * 1e-3f, b = 0.5f + threadIdx.x * 1e-3f;
  for (int j = 0; j < 32; ++j) c[j] = j;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 32; ++j) c[j] = fmaf(a, c[j], b);
  }
  float s = 0.f;
  for (int j = 0; j < 32; ++j) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int threads : {256, 512, 1024}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      mma_kernel<8><<<148, threads>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double flops = 2.0 * 16 * 8 * 8 * 8.0 * iters * (threads / 32) * 148;
      if (rep) printf("mma.sync m16n8k8 tf32: %4d threads/SM  %.3f ms  %.1f TFLOP/s (tf32 dense)\n", threads, ms, flops / ms / 1e9);
    }
  }
  for (int threads : {256, 512, 1024}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      ffma_kernel<<<148, threads>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double flops = 2.0 * 32.0 * iters * threads * 148;
      if (rep) printf("ffma: %4d threads/SM  %.3f ms  %.1f TFLOP/s\n", threads, ms, flops / ms / 1e9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
// Measured on B200 (gpurun, 2026-10-18): mma.sync m16n8k8 tf32 278 TFLOP/s dense at 8 / 16 / 32 warps per SM,
// FFMA 72 TFLOP/s.
