// common.cuh -- shared host/device helpers for the cgl_b200 engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/cgl_b200.h"

namespace cgl {

// ---- error reporting (thread-local message, C return codes) -------------------------------
void set_error(const char* fmt, ...);
// every kernel launch of the library is counted (cgl_launch_count: bench.py's gpu_launches)
void count_launch(int n = 1);
// optional per-kernel-class timing with CUDA events on the launching stream (cgl_profile_*): bench.py's
// roofline numbers are measured with these, live, inside the timed region
void prof_begin(int tag, double bytes, double flops, cudaStream_t st);
void prof_end(cudaStream_t st);
struct ProfScope {
  cudaStream_t st;
  ProfScope(int tag, double bytes, double flops, cudaStream_t s) : st(s) { prof_begin(tag, bytes, flops, s); }
  ~ProfScope() { prof_end(st); }
};

#define CGL_CHECK_CUDA(expr)                                                          \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      cgl::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return CGL_ECUDA;                                                               \
    }                                                                                 \
  } while (0)

#define CGL_CHECK_LAUNCH()               \
  do {                                   \
    cgl::count_launch();                 \
    CGL_CHECK_CUDA(cudaGetLastError());  \
  } while (0)

#define CGL_REQUIRE(cond, ...)      \
  do {                              \
    if (!(cond)) {                  \
      cgl::set_error(__VA_ARGS__);  \
      return CGL_EINVAL;            \
    }                               \
  } while (0)

// cudaFuncSetAttribute is per device: "already done" flags are bit masks over the device ordinal
static inline bool first_use_on_device(unsigned long long& mask) {
  int d = 0;
  cudaGetDevice(&d);
  const unsigned long long b = 1ull << (d & 63);
  if (mask & b) return false;
  mask |= b;
  return true;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- Adam (torch.optim.Adam, single-tensor path; reference a7: CGLGAN/2DMG/main.py:192,337) --
// Scalars that torch computes on the host in double for step t and casts to the tensor dtype.
struct AdamScalars {
  float lerp_coeff;    // lerp_(g, 1-b1): coeff = w (w<0.5) or w-1
  int lerp_small;      // |w| < 0.5
  float beta2;         // v.mul_(beta2)
  float one_minus_b2;  // addcmul_(g, g, value=1-beta2)
  float bc2_sqrt;      // sqrt(1 - beta2^t)
  float neg_step_size; // -(lr / (1 - beta1^t))
  float eps;
  float inv_bc2_sqrt;  // 1 / sqrt(1 - beta2^t)   (adam_update_fast)
};

__host__ __device__ inline AdamScalars make_adam_scalars(int t, float lr, float b1, float b2, float eps) {
  AdamScalars s;
  // torch keeps lr/betas as python floats (double); the fp32 knobs are widened exactly.
  double w = 1.0 - (double)b1;
  s.lerp_small = (fabs(w) < 0.5) ? 1 : 0;
  s.lerp_coeff = (float)(s.lerp_small ? w : (w - 1.0));
  s.beta2 = b2;
  s.one_minus_b2 = (float)(1.0 - (double)b2);
  double bc1 = 1.0 - pow((double)b1, (double)t);
  double bc2 = 1.0 - pow((double)b2, (double)t);
  s.bc2_sqrt = (float)sqrt(bc2);
  s.neg_step_size = (float)(-((double)lr / bc1));
  s.eps = eps;
  s.inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
  return s;
}

// One element of: m.lerp_(g,1-b1); v.mul_(b2).addcmul_(g,g,1-b2);
//                 denom = v.sqrt()/bc2_sqrt + eps; p.addcdiv_(m, denom, -step_size)
// Every product/sum is rounded separately, in the order ATen's CPU kernels apply them.
__device__ __forceinline__ void adam_update(float& p, float& m, float& v, float g, const AdamScalars& s) {
  float diff = __fsub_rn(g, m);
  float base = s.lerp_small ? m : g;
  m = __fmaf_rn(s.lerp_coeff, diff, base);
  float v1 = __fmul_rn(v, s.beta2);
  v = __fadd_rn(v1, __fmul_rn(__fmul_rn(s.one_minus_b2, g), g));
  float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), s.bc2_sqrt), s.eps);
  p = __fadd_rn(p, __fdiv_rn(__fmul_rn(s.neg_step_size, m), denom));
}

// The same step with the two IEEE divisions and the IEEE square root replaced by MUFU approximations
// (sqrt.approx / rcp.approx: <= 1 ulp each; the division by bc2_sqrt becomes a multiplication by its reciprocal).
// m and v are still bit-identical to torch; the INCREMENT of p carries a relative error <= ~4e-7, i.e. below
// 2e-10 absolute at lr = 2e-4 -- a tenth of an ulp of a weight of 0.03, against a parity bar of 1e-5. Used by
// the tcgen05 weight-gradient epilogue only, which ncu shows to be bound by instruction issue (the exact
// version is ~48 instructions per element with two FCHK/branch slow paths, this one ~16).
__device__ __forceinline__ void adam_update_fast(float& p, float& m, float& v, float g, const AdamScalars& s) {
  float diff = __fsub_rn(g, m);
  float base = s.lerp_small ? m : g;
  m = __fmaf_rn(s.lerp_coeff, diff, base);
  float v1 = __fmul_rn(v, s.beta2);
  v = __fadd_rn(v1, __fmul_rn(__fmul_rn(s.one_minus_b2, g), g));
  float sq, r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(sq) : "f"(v));
  const float denom = __fmaf_rn(sq, s.inv_bc2_sqrt, s.eps);
  asm("rcp.approx.f32 %0, %1;" : "=f"(r) : "f"(denom));
  p = __fmaf_rn(__fmul_rn(s.neg_step_size, m), r, p);
}

// ---- activations ---------------------------------------------------------------------------
__device__ __forceinline__ float act_fwd(float x, int act, float slope) {
  switch (act) {
    case CGL_ACT_LRELU: return x > 0.f ? x : x * slope;
    case CGL_ACT_TANH: return tanhf(x);
    case CGL_ACT_SIGMOID: return 1.f / (1.f + expf(-x));
    default: return x;
  }
}
// derivative expressed through the saved OUTPUT a = act(x)
__device__ __forceinline__ float act_bwd_from_out(float a, int act, float slope) {
  switch (act) {
    case CGL_ACT_LRELU: return a > 0.f ? 1.f : slope;
    case CGL_ACT_TANH: return 1.f - a * a;
    case CGL_ACT_SIGMOID: return a * (1.f - a);
    default: return 1.f;
  }
}

}  // namespace cgl
