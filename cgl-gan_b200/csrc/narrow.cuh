// narrow.cuh -- the three Linear products for NARROW layers (in * out <= NARROW_MAX_W weights: the 2 -> 128 input
// layer of the 2DMG discriminator, the 100 -> 32 trunk and the 32 -> 2 heads of the 2DMG generator,
// CGLGAN/2DMG/model.py:26-71). On the 128 x 128 FFMA tiles of gemm.cuh these layers are mostly padding (K = 2, or N = 2)
// and cost 60-150 us per launch for a few MB of traffic. Here one CTA per group keeps the layer's whole weight matrix
// in shared memory -- the part of the north-star design that does fit: small weights resident on chip -- and every
// thread produces output elements with a plain fp32 FMA chain in ascending contraction order (exact-fp32 semantics,
// like the FFMA GEMM). The kernels stream the wide side (activations) once: HBM-bound.
#pragma once
#include "gemm.cuh"

namespace cgl {

constexpr int NARROW_MAX_W = 4096;    // weights of a layer that may use these kernels (16 KB of shared memory)
constexpr int NARROW_THREADS = 256;

// y[g][r][o] = act( sum_i x[g][r][i] * W[g][o][i] + b[g][o] )
static __global__ void __launch_bounds__(NARROW_THREADS) narrow_fwd_kernel(int rows, int in, int out, RowMap X, const float* params,
                                                                   long long ldp, const int* ids, long long w_off,
                                                                   long long b_off, int act, float slope, float* y,
                                                                   long long y_gstride) {
  extern __shared__ float nsm[];
  float* sW = nsm;               // [in][out]: transposed, so that consecutive lanes (consecutive o) hit consecutive banks
  float* sb = sW + out * in;     // [out]
  const int g = blockIdx.x;
  const int rowid = ids ? ids[g] : g;
  const float* W = params + (long long)rowid * ldp + w_off;
  for (int i = threadIdx.x; i < out * in; i += NARROW_THREADS) {
    const int o = i / in, k = i - o * in;
    sW[k * out + o] = W[i];
  }
  for (int i = threadIdx.x; i < out; i += NARROW_THREADS) sb[i] = (b_off >= 0) ? params[(long long)rowid * ldp + b_off + i] : 0.f;
  __syncthreads();
  const Rows R = resolve(X, g);
  float* yg = y + (long long)g * y_gstride;
  // consecutive threads -> consecutive o of one row: coalesced stores, the row's inputs are broadcast loads
  for (int e = blockIdx.y * NARROW_THREADS + threadIdx.x; e < rows * out; e += gridDim.y * NARROW_THREADS) {
    const int r = e / out, o = e - r * out;
    const float* xr = row_ptr(R, r);
    const float* w = sW + o;
    float acc = 0.f;
#pragma unroll 4
    for (int i = 0; i < in; ++i) acc = fmaf(__ldg(xr + i), w[i * out], acc);
    yg[e] = act_fwd(acc + sb[o], act, slope);
  }
}

// dx[g][r][i] = ( sum_o dy[g][r][o] * W[g][o][i] ) * act'(saved[g][r][i])
static __global__ void __launch_bounds__(NARROW_THREADS) narrow_bwd_data_kernel(int rows, int in, int out, const float* dy,
                                                                        long long dy_gstride, const float* params,
                                                                        long long ldp, const int* ids, long long w_off,
                                                                        const float* saved, long long saved_gstride, int act,
                                                                        float slope, float* dx, long long dx_gstride) {
  extern __shared__ float nsm[];
  float* sW = nsm;               // [out][in]
  const int g = blockIdx.x;
  const int rowid = ids ? ids[g] : g;
  const float* W = params + (long long)rowid * ldp + w_off;
  for (int i = threadIdx.x; i < out * in; i += NARROW_THREADS) sW[i] = W[i];
  __syncthreads();
  const float* dyg = dy + (long long)g * dy_gstride;
  float* dxg = dx + (long long)g * dx_gstride;
  const float* sg = saved ? saved + (long long)g * saved_gstride : nullptr;
  for (int e = blockIdx.y * NARROW_THREADS + threadIdx.x; e < rows * in; e += gridDim.y * NARROW_THREADS) {
    const int r = e / in, i = e - r * in;
    const float* d = dyg + (long long)r * out;
    float acc = 0.f;
#pragma unroll 4
    for (int o = 0; o < out; ++o) acc = fmaf(__ldg(d + o), sW[o * in + i], acc);
    if (sg) acc *= act_bwd_from_out(sg[e], act, slope);
    dxg[e] = acc;
  }
}

// dW[g][o][i] = sum_r dy[g][r][o] * x[g][r][i],  db[g][o] = sum_r dy[g][r][o]; Adam step (ADAM) or plain store.
// One thread per weight (and per bias), rows ascending.
template <bool ADAM>
__global__ void __launch_bounds__(NARROW_THREADS) narrow_wgrad_kernel(int rows, int in, int out, const float* dy,
                                                                     long long dy_gstride, RowMap X, float* base,
                                                                     long long ld, const int* ids, long long w_off,
                                                                     long long b_off, float* adam_m, float* adam_v,
                                                                     const int* step, float lr, float b1, float b2, float eps,
                                                                     const AdamScalars* scal) {
  const int g = blockIdx.x;
  const int rowid = ids ? ids[g] : g;
  const Rows R = resolve(X, g);
  const float* dyg = dy + (long long)g * dy_gstride;
  AdamScalars s = {};
  if (ADAM) s = scal ? scal[g] : make_adam_scalars(step[rowid], lr, b1, b2, eps);
  const int nw = out * in;
  const int total = nw + (b_off >= 0 ? out : 0);
  for (int e = threadIdx.x; e < total; e += NARROW_THREADS) {
    float acc = 0.f;
    long long off;
    if (e < nw) {
      const int o = e / in, i = e - o * in;
#pragma unroll 4
      for (int r = 0; r < rows; ++r) acc = fmaf(__ldg(dyg + (long long)r * out + o), __ldg(row_ptr(R, r) + i), acc);
      off = (long long)rowid * ld + w_off + e;
    } else {
      const int o = e - nw;
      for (int r = 0; r < rows; ++r) acc += __ldg(dyg + (long long)r * out + o);
      off = (long long)rowid * ld + b_off + o;
    }
    if (ADAM) {
      float w = base[off], mm = adam_m[off], vv = adam_v[off];
      adam_update(w, mm, vv, acc, s);
      base[off] = w; adam_m[off] = mm; adam_v[off] = vv;
    } else {
      base[off] = acc;
    }
  }
}

static inline bool narrow_ok(int in, int out) { return (long long)in * out <= NARROW_MAX_W; }
static inline unsigned narrow_grid_y(int elems) {
  const int per = (elems + NARROW_THREADS - 1) / NARROW_THREADS;
  return (unsigned)(per < 1 ? 1 : (per > 8 ? 8 : per));   // up to 8 CTAs share a group's rows (each reloads 16 KB of W)
}

}  // namespace cgl
