// builders.cuh -- GemmParams for the three Linear products (forward, data gradient, weight gradient).
#pragma once
#include "gemm.cuh"

namespace cgl {

// ---- GemmParams builders ------------------------------------------------------------------
// y[g][r][o] = act( sum_i x[g][r][i] * W[g][o][i] + b[g][o] )
static inline GemmParams fwd_params(int rows, int in, int out, const RowMap& X, const float* params, long long ldp,
                             const int* ids, long long w_off, long long b_off, int act, float slope, float* y,
                             long long y_gstride) {
  GemmParams p = {};
  p.M = rows; p.N = out; p.K = in;
  p.A = X;
  p.B = single_rows(params + w_off, ldp, ids, in);
  p.cbase = y; p.c_gstride = y_gstride; p.cidx = nullptr; p.c_off = 0; p.ldc = out;
  p.c_vec = (aligned16(y) && out % 4 == 0 && y_gstride % 4 == 0) ? 1 : 0;
  p.bias_base = (b_off >= 0) ? params : nullptr; p.bias_gstride = ldp; p.bias_idx = ids; p.bias_off = b_off;
  p.act = act; p.slope = slope;
  p.dbias_off = -1;
  return p;
}
// dx[g][r][i] = ( sum_o dy[g][r][o] * W[g][o][i] ) * act'(saved[g][r][i])
static inline GemmParams bwd_data_params(int rows, int in, int out, const float* dy, long long dy_gstride,
                                  const float* params, long long ldp, const int* ids, long long w_off,
                                  const float* saved, long long saved_gstride, int act, float slope, float* dx,
                                  long long dx_gstride) {
  GemmParams p = {};
  p.M = rows; p.N = in; p.K = out;
  p.A = single_rows(dy, dy_gstride, nullptr, out);
  p.B = single_rows(params + w_off, ldp, ids, in);  // row = contraction index o, contiguous along i
  p.cbase = dx; p.c_gstride = dx_gstride; p.cidx = nullptr; p.c_off = 0; p.ldc = in;
  p.c_vec = (aligned16(dx) && in % 4 == 0 && dx_gstride % 4 == 0) ? 1 : 0;
  p.saved = saved; p.saved_gstride = saved_gstride;
  p.act = act; p.slope = slope;
  p.dbias_off = -1;
  return p;
}
// dW[g][o][i] = sum_r dy[g][r][o] * x[g][r][i]   (+ db[g][o] = sum_r dy[g][r][o])
static inline GemmParams wgrad_params(int rows, int in, int out, const float* dy, long long dy_gstride, const RowMap& X,
                               float* base, long long ld, const int* ids, long long w_off, long long b_off) {
  GemmParams p = {};
  p.M = out; p.N = in; p.K = rows;
  p.A = single_rows(dy, dy_gstride, nullptr, out);  // row = contraction index r, contiguous along o
  p.B = X;                                           // row = contraction index r, contiguous along i
  p.cbase = base; p.c_gstride = ld; p.cidx = ids; p.c_off = w_off; p.ldc = in;
  p.c_vec = (aligned16(base) && ld % 4 == 0 && w_off % 4 == 0 && in % 4 == 0) ? 1 : 0;
  p.bias_off = b_off; p.dbias_off = b_off;
  return p;
}


}  // namespace cgl
