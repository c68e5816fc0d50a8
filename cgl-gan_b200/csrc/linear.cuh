// linear.cuh -- the three products of a grouped Linear layer (forward, data gradient, weight gradient
// [+ fused Adam]) dispatched onto one of the two GEMM kernels of this engine:
//   tc_gemm.cuh : tcgen05 / TMEM 3xTF32 grouped GEMM, for the layers that are real dense contractions
//                 (the MNIST 784/512/256/1024-wide layers)
//   gemm.cuh    : exact-fp32 FFMA grouped GEMM, for everything narrow or unaligned (the 2-wide 2DMG input,
//                 ragged pointers)
// Both are sm_100a kernels of this library; the choice is by shape/alignment only (cgl_set_gemm_mode
// can pin one for tests and profiling).
#pragma once
#include "builders.cuh"
#include "narrow.cuh"
#include "tc_persist.cuh"
#include "tc_pair.cuh"
#include "tc_sweep.cuh"
#include "tc_tma.cuh"
#include "tc_tma_persist.cuh"

namespace cgl {

enum { GEMM_AUTO = 0, GEMM_FFMA = 1, GEMM_TC = 2 };
// narrow layers leave the GEMM kernels in the automatic mode only (GEMM_FFMA / GEMM_TC pin one GEMM kernel for tests)
static inline bool narrow_wanted(int in, int out);
int gemm_mode();  // defined in dstep.cu

// db[g][o] = sum_r dy[g][r][o], then either stored or applied as an Adam step on the bias.
template <bool ADAM>
__global__ void __launch_bounds__(128) bias_grad_kernel(int rows, int out, const float* __restrict__ dy,
                                                        long long dy_gstride, float* base, long long ld,
                                                        const int* ids, long long b_off, float* adam_m,
                                                        float* adam_v, const int* step, float lr, float b1, float b2,
                                                        float eps, const AdamScalars* scal) {
  const int g = blockIdx.y;
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= out) return;
  const float* d = dy + (long long)g * dy_gstride + o;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int r = 0;
  for (; r + 3 < rows; r += 4) {
    s0 += d[(long long)r * out];
    s1 += d[(long long)(r + 1) * out];
    s2 += d[(long long)(r + 2) * out];
    s3 += d[(long long)(r + 3) * out];
  }
  for (; r < rows; ++r) s0 += d[(long long)r * out];
  const float gsum = (s0 + s1) + (s2 + s3);
  const int rowid = ids ? ids[g] : g;
  const long long off = (long long)rowid * ld + b_off + o;
  if (ADAM) {
    const AdamScalars s = scal ? scal[g] : make_adam_scalars(step[rowid], lr, b1, b2, eps);
    float w = base[off], mm = adam_m[off], vv = adam_v[off];
    adam_update(w, mm, vv, gsum, s);
    base[off] = w; adam_m[off] = mm; adam_v[off] = vv;
  } else {
    base[off] = gsum;
  }
}

static inline bool narrow_wanted(int in, int out) { return gemm_mode() == GEMM_AUTO && narrow_ok(in, out); }

static inline bool tc_wanted(int m_dim, int k_dim, bool operands_ok) {
  const int mode = gemm_mode();
  if (mode == GEMM_FFMA) return false;
  if (!operands_ok) return false;
  if (mode == GEMM_TC) return true;
  return m_dim >= 64 && k_dim >= 32;
}

// y[g][r][o] = act( sum_i x[g][r][i] * W[g][o][i] + b[g][o] )
static inline cudaError_t run_linear_fwd(int G, int rows, int in, int out, const RowMap& X, const float* params,
                                         long long ldp, const int* ids, long long w_off, long long b_off, int act,
                                         float slope, float* y, long long y_gstride, cudaStream_t st) {
  RowMap W = single_rows(params + w_off, ldp, ids, in);
  const bool tc = tc_wanted(out, in, tc_rowmap_ok(W, in) && tc_rowmap_ok(X, in));
  // algorithmic traffic: W, b, x read once, y written once
  ProfScope prof(tc ? CGL_PROF_FWD_TC : CGL_PROF_FWD_FFMA,
                 4.0 * G * ((double)out * in + out + (double)rows * in + (double)rows * out), 2.0 * G * rows * (double)in * out, st);
  if (!tc && narrow_wanted(in, out)) {
    dim3 grid(G, narrow_grid_y(rows * out));
    narrow_fwd_kernel<<<grid, NARROW_THREADS, (size_t)(out * in + out) * sizeof(float), st>>>(
        rows, in, out, X, params, ldp, ids, w_off, b_off, act, slope, y, y_gstride);
    count_launch();
    return cudaGetLastError();
  }
  if (tc) {
    TcParams p = {};
    p.M = out; p.N = rows; p.K = in;
    p.A = W; p.B = X;
    p.cbase = y; p.c_gstride = y_gstride; p.cidx = nullptr; p.c_off = 0; p.ldc = out;
    p.c_vec = (aligned16(y) && y_gstride % 4 == 0 && out % 4 == 0 &&
               (b_off < 0 || (aligned16(params) && ldp % 4 == 0 && b_off % 4 == 0))) ? 1 : 0;
    p.bias_base = (b_off >= 0) ? params : nullptr; p.bias_gstride = ldp; p.bias_idx = ids; p.bias_off = b_off;
    p.act = act; p.slope = slope;
    cudaError_t pe = cudaSuccess;
    if (launch_tc_tma_persistent<true, EPI_FWD>(p, G, st, &pe)) return pe;
    if (launch_tc_tma<true, EPI_FWD>(p, G, st, &pe)) return pe;
    if (launch_tc_sweep<true, true, EPI_FWD>(p, G, st, &pe)) return pe;
    if ((tc_tune() & 256) && launch_tc_pair<true, EPI_FWD>(p, G, st, &pe)) return pe;
    if (launch_tc_persistent<true, true, EPI_FWD>(p, G, st, &pe)) return pe;
    return launch_tc_gemm<true, true, EPI_FWD>(p, G, st);
  }
  GemmParams p = fwd_params(rows, in, out, X, params, ldp, ids, w_off, b_off, act, slope, y, y_gstride);
  return launch_grouped_gemm<true, true, EPI_FWD>(p, G, st);
}

// dx[g][r][i] = ( sum_o dy[g][r][o] * W[g][o][i] ) * act'(saved[g][r][i])     (saved NULL: plain product)
static inline cudaError_t run_linear_bwd_data(int G, int rows, int in, int out, const float* dy, long long dy_gstride,
                                              const float* params, long long ldp, const int* ids, long long w_off,
                                              const float* saved, long long saved_gstride, int act, float slope,
                                              float* dx, long long dx_gstride, cudaStream_t st) {
  RowMap W = single_rows(params + w_off, ldp, ids, in);
  RowMap DY = single_rows(dy, dy_gstride, nullptr, out);
  const bool tc = tc_wanted(in, out, tc_rowmap_ok(W, in) && tc_rowmap_ok(DY, out));
  ProfScope prof(tc ? CGL_PROF_BWD_TC : CGL_PROF_BWD_FFMA,
                 4.0 * G * ((double)out * in + (double)rows * out + (saved ? 2.0 : 1.0) * rows * in),
                 2.0 * G * rows * (double)in * out, st);
  if (!tc && narrow_wanted(in, out)) {
    dim3 grid(G, narrow_grid_y(rows * in));
    narrow_bwd_data_kernel<<<grid, NARROW_THREADS, (size_t)out * in * sizeof(float), st>>>(
        rows, in, out, dy, dy_gstride, params, ldp, ids, w_off, saved, saved_gstride, act, slope, dx, dx_gstride);
    count_launch();
    return cudaGetLastError();
  }
  if (tc) {
    TcParams p = {};
    p.M = in; p.N = rows; p.K = out;
    p.A = W;   // MN-major: line = o (contraction), contiguous along i
    p.B = DY;  // K-major : line = r, contiguous along o
    p.cbase = dx; p.c_gstride = dx_gstride; p.cidx = nullptr; p.c_off = 0; p.ldc = in;
    p.c_vec = (aligned16(dx) && dx_gstride % 4 == 0 && in % 4 == 0 &&
               (!saved || (aligned16(saved) && saved_gstride % 4 == 0))) ? 1 : 0;
    p.saved = saved; p.saved_gstride = saved_gstride; p.act = act; p.slope = slope;
    cudaError_t pe = cudaSuccess;
    if (launch_tc_tma_persistent<false, EPI_BWD_DATA>(p, G, st, &pe)) return pe;
    if (launch_tc_tma<false, EPI_BWD_DATA>(p, G, st, &pe)) return pe;
    if (launch_tc_sweep<false, true, EPI_BWD_DATA>(p, G, st, &pe)) return pe;                  // saved == NULL: plain store
    if ((tc_tune() & 512) && launch_tc_pair<false, EPI_BWD_DATA>(p, G, st, &pe)) return pe;
    if (saved) {
      if (launch_tc_persistent<false, true, EPI_BWD_DATA>(p, G, st, &pe)) return pe;
      return launch_tc_gemm<false, true, EPI_BWD_DATA>(p, G, st);
    }
    if (launch_tc_persistent<false, true, EPI_STORE>(p, G, st, &pe)) return pe;
    return launch_tc_gemm<false, true, EPI_STORE>(p, G, st);
  }
  GemmParams p = bwd_data_params(rows, in, out, dy, dy_gstride, params, ldp, ids, w_off, saved, saved_gstride, act,
                                 slope, dx, dx_gstride);
  if (saved) return launch_grouped_gemm<true, false, EPI_BWD_DATA>(p, G, st);
  return launch_grouped_gemm<true, false, EPI_STORE>(p, G, st);
}

struct AdamArgs {
  float* m; float* v; const int* step; float lr, b1, b2, eps;
  const AdamScalars* scal;  // [G] per-group scalars from adam_prepare_kernel (NULL: derived in every kernel)
};

// step[row(g)] += 1 and the step's Adam scalars (bias corrections: two double-precision pow) computed ONCE per
// group; every fused-Adam epilogue of the call then reads 7 floats instead of redoing fp64 math per thread.
static __global__ void adam_prepare_kernel(int G, int* step, const int* ids, float lr, float b1, float b2, float eps,
                                           AdamScalars* scal) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  const int row = ids ? ids[g] : g;
  const int t = step[row] + 1;
  step[row] = t;
  scal[g] = make_adam_scalars(t, lr, b1, b2, eps);
}
// the scalars alone, for a step counter that already includes the update (cgl_linear_wgrad_adam)
static __global__ void adam_scalars_kernel(int G, const int* step, const int* ids, float lr, float b1, float b2,
                                           float eps, AdamScalars* scal) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= G) return;
  scal[g] = make_adam_scalars(step[ids ? ids[g] : g], lr, b1, b2, eps);
}

// dW[g][o][i] = sum_r dy[g][r][o] * x[g][r][i], db[g][o] = sum_r dy[g][r][o];
// adam != NULL: applied in the epilogue as an Adam step on W / b (base = params);
// adam == NULL: stored at base (a gradient buffer with the packed-row layout). b_off < 0: no bias.
static inline cudaError_t run_linear_wgrad(int G, int rows, int in, int out, const float* dy, long long dy_gstride,
                                           const RowMap& X, float* base, long long ld, const int* ids, long long w_off,
                                           long long b_off, const AdamArgs* adam, cudaStream_t st) {
  RowMap DY = single_rows(dy, dy_gstride, nullptr, out);
  const bool tc = tc_wanted(in, rows, tc_rowmap_ok(X, in) && tc_rowmap_ok(DY, out));
  // Adam: W, m, v (and the bias triple) read and written = 24 B per parameter; x and dy read once
  const double nparam = (double)out * in + (b_off >= 0 ? out : 0);
  ProfScope prof(adam ? (tc ? CGL_PROF_WGRAD_ADAM_TC : CGL_PROF_WGRAD_ADAM_FFMA) : (tc ? CGL_PROF_WGRAD_TC : CGL_PROF_WGRAD_FFMA),
                 G * ((adam ? 24.0 : 4.0) * nparam + 4.0 * rows * ((double)in + out)), 2.0 * G * rows * (double)in * out, st);
  if (!tc && narrow_wanted(in, out)) {
    if (adam)
      narrow_wgrad_kernel<true><<<G, NARROW_THREADS, 0, st>>>(rows, in, out, dy, dy_gstride, X, base, ld, ids, w_off, b_off,
                                                             adam->m, adam->v, adam->step, adam->lr, adam->b1, adam->b2,
                                                             adam->eps, adam->scal);
    else
      narrow_wgrad_kernel<false><<<G, NARROW_THREADS, 0, st>>>(rows, in, out, dy, dy_gstride, X, base, ld, ids, w_off, b_off,
                                                              nullptr, nullptr, nullptr, 0.f, 0.f, 0.f, 0.f, nullptr);
    count_launch();
    return cudaGetLastError();
  }
  if (tc) {
    TcParams p = {};
    p.M = in; p.N = out; p.K = rows;
    p.A = X;   // MN-major: line = r, contiguous along i
    p.B = DY;  // MN-major: line = r, contiguous along o
    p.cbase = base; p.c_gstride = ld; p.cidx = ids; p.c_off = w_off; p.ldc = in;
    p.c_vec = (aligned16(base) && ld % 4 == 0 && w_off % 4 == 0 && in % 4 == 0 &&
               (!adam || (aligned16(adam->m) && aligned16(adam->v)))) ? 1 : 0;
    cudaError_t e;
    if (adam) {
      p.adam_m = adam->m; p.adam_v = adam->v; p.step = adam->step; p.scal = adam->scal;
      p.lr = adam->lr; p.b1 = adam->b1; p.b2 = adam->b2; p.eps = adam->eps;
      e = launch_tc_gemm<false, false, EPI_ADAM>(p, G, st);
    } else {
      e = launch_tc_gemm<false, false, EPI_STORE>(p, G, st);
    }
    if (e != cudaSuccess || b_off < 0) return e;
    dim3 grid((out + 127) / 128, G);
    if (adam) {
      bias_grad_kernel<true><<<grid, 128, 0, st>>>(rows, out, dy, dy_gstride, base, ld, ids, b_off, adam->m, adam->v,
                                                   adam->step, adam->lr, adam->b1, adam->b2, adam->eps, adam->scal);
    } else {
      bias_grad_kernel<false><<<grid, 128, 0, st>>>(rows, out, dy, dy_gstride, base, ld, ids, b_off, nullptr, nullptr,
                                                    nullptr, 0.f, 0.f, 0.f, 0.f, nullptr);
    }
    count_launch();
    return cudaGetLastError();
  }
  GemmParams p = wgrad_params(rows, in, out, dy, dy_gstride, X, base, ld, ids, w_off, b_off);
  if (b_off < 0) { p.bias_off = -1; p.dbias_off = -1; }
  if (adam) {
    p.adam_m = adam->m; p.adam_v = adam->v; p.step = adam->step; p.scal = adam->scal;
    p.lr = adam->lr; p.b1 = adam->b1; p.b2 = adam->b2; p.eps = adam->eps;
    return launch_grouped_gemm<false, false, EPI_ADAM>(p, G, st);
  }
  return launch_grouped_gemm<false, false, EPI_STORE>(p, G, st);
}

}  // namespace cgl
