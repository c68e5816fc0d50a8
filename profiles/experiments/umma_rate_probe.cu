// umma_rate_probe.cu -- what ONE tcgen05.mma costs on sm_100a, per kind / N / operand source / accumulator pattern:
// one CTA per SM, one warp issues REP instructions back to back (operands: zeroed shared memory / TMEM garbage), one commit,
// clock64 around the whole batch.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o umma_rate_probe umma_rate_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
// KIND 0: tf32 (K = 8), 1: bf16 (K = 16). AT: A from TMEM. All lanes call; the elected lane issues.
template <int KIND, bool AT>
__device__ __forceinline__ void mma(uint32_t d, uint32_t a_t, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (AT) {
    if (KIND == 0)
      asm volatile("{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_t), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    else
      asm volatile("{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_t), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  } else {
    if (KIND == 0)
      asm volatile("{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    else
      asm volatile("{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
  }
}

template <int KIND, bool AT>
__global__ void __launch_bounds__(128, 1) probe(int N, int rep, int n_acc, long long* out) {
  extern __shared__ __align__(1024) char sm[];
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t slot;
  char* s = sm + ((1024u - (smem_u32(sm) & 1023u)) & 1023u);
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<float*>(s)[i] = 0.f;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = slot;
  if (warp == 0) {
    const uint32_t fmt = KIND == 0 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t ad = umma_desc(smem_u32(s), 16u, 1024u, 2u);
    const uint64_t bd = umma_desc(smem_u32(s) + 16384u, 16u, 1024u, 2u);
    const uint32_t a_t = tm + (uint32_t)(n_acc * N);
    const long long t0 = clock64();
    int r = 0;
    for (int i = 0; i < rep; ++i) {
      mma<KIND, AT>(tm + (uint32_t)(r * N), a_t, ad + (uint64_t)(2 * (i & 3)), bd + (uint64_t)(2 * (i & 3)), idesc, i >= n_acc ? 1u : 0u);
      if (++r == n_acc) r = 0;
    }
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    const long long t1 = clock64();
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    const long long t2 = clock64();
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

template <int KIND, bool AT>
void run(const char* name, int N, int n_acc, int ctas) {
  const int rep = 2048;
  long long* d;
  cudaMalloc(&d, sizeof(long long) * 2 * ctas);
  cudaFuncSetAttribute(probe<KIND, AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 3; ++w) probe<KIND, AT><<<ctas, 128, 60 * 1024>>>(N, rep, n_acc, d);
  cudaEventRecord(a);
  probe<KIND, AT><<<ctas, 128, 60 * 1024>>>(N, rep, n_acc, d);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const double k = KIND == 0 ? 8.0 : 16.0;
  printf("%-28s N=%3d acc-regions=%d ctas=%3d: issue %6.1f clk/MMA, complete %6.1f clk/MMA, %.1f ns/MMA, %.0f TFLOP/s (%s)\n", name, N, n_acc, ctas,
         (double)h[0] / rep, (double)h[1] / rep, ms * 1e6 / rep, 2.0 * 128 * N * k * rep * ctas / (ms * 1e-3) / 1e12, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int ctas : {1, 148}) {
    for (int N : {64, 112, 128, 256}) {
      run<0, false>("tf32 A smem", N, 1, ctas);
      run<0, true>("tf32 A tmem", N, 1, ctas);
      run<1, false>("bf16 A smem", N, 1, ctas);
      run<1, true>("bf16 A tmem", N, 1, ctas);
    }
    run<0, true>("tf32 A tmem, 3 regions", 112, 3, ctas);
    run<0, false>("tf32 A smem, 3 regions", 112, 3, ctas);
    run<0, true>("tf32 A tmem, 4 regions", 64, 4, ctas);
  }
  return 0;
}
