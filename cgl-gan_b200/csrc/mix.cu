// mix.cu -- K3: aggregation over packed parameter rows (pure HBM streaming), the fused Adam over
// packed rows used by the server-side generator, and the per-server reduction of dLoss/dXg.
// Reference: Cloud.run CGLGAN/2DMG/main.py:116-136; FL Server.run FLGAN/MNIST/flgan.py:143-163;
// segema mix CGLGAN/2DMG/main.py:205-208; MD-GAN swap MDGAN/MNIST/mdgan.py:158-164.
#include "common.cuh"

namespace cgl {

constexpr int MIX_THREADS = 256;

__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// acc = acc + w*x with product and sum rounded separately (the `p[key] += paras[key] * A` of
// Cloud.run is a mul kernel followed by an add kernel).
__device__ __forceinline__ void axpy4(float4& acc, float w, const float4& x) {
  acc.x = __fadd_rn(acc.x, __fmul_rn(x.x, w));
  acc.y = __fadd_rn(acc.y, __fmul_rn(x.y, w));
  acc.z = __fadd_rn(acc.z, __fmul_rn(x.z, w));
  acc.w = __fadd_rn(acc.w, __fmul_rn(x.w, w));
}

// dst[r,:] = sum_j vals[j] * src[col[j],:]; the first term initialises (p[key] = paras*A).
// MEAN: dst[r,:] = (sum_j src[col[j],:]) / count -- receive_parameter's `p += d ... p /= len` (CGLGAN/2DMG/main.py:171-179).
__device__ __forceinline__ float4 mul4(const float4& x, float w) {
  return make_float4(__fmul_rn(x.x, w), __fmul_rn(x.y, w), __fmul_rn(x.z, w), __fmul_rn(x.w, w));
}
template <bool VEC, bool MEAN>
__global__ void __launch_bounds__(MIX_THREADS) mix_csr_kernel(long long n, const int* __restrict__ row_ptr,
                                                             const int* __restrict__ col,
                                                             const float* __restrict__ vals,
                                                             const float* __restrict__ src, long long ld_src,
                                                             float* __restrict__ dst, long long ld_dst) {
  const int r = blockIdx.y;
  const int j0 = row_ptr[r], j1 = row_ptr[r + 1];
  const float cnt = (float)(j1 - j0);
  if (VEC) {
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * MIX_THREADS + threadIdx.x; i < n4;
         i += (long long)gridDim.x * MIX_THREADS) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      int j = j0;
      if (j < j1) {
        float4 x = ld_stream(reinterpret_cast<const float4*>(src + (long long)col[j] * ld_src) + i);
        acc = MEAN ? x : mul4(x, vals[j]);
        ++j;
      }
      for (; j + 3 < j1; j += 4) {
        float4 x0 = ld_stream(reinterpret_cast<const float4*>(src + (long long)col[j] * ld_src) + i);
        float4 x1 = ld_stream(reinterpret_cast<const float4*>(src + (long long)col[j + 1] * ld_src) + i);
        float4 x2 = ld_stream(reinterpret_cast<const float4*>(src + (long long)col[j + 2] * ld_src) + i);
        float4 x3 = ld_stream(reinterpret_cast<const float4*>(src + (long long)col[j + 3] * ld_src) + i);
        axpy4(acc, MEAN ? 1.f : vals[j], x0); axpy4(acc, MEAN ? 1.f : vals[j + 1], x1);     // x * 1 is exact
        axpy4(acc, MEAN ? 1.f : vals[j + 2], x2); axpy4(acc, MEAN ? 1.f : vals[j + 3], x3);
      }
      for (; j < j1; ++j) {
        float4 x = ld_stream(reinterpret_cast<const float4*>(src + (long long)col[j] * ld_src) + i);
        axpy4(acc, MEAN ? 1.f : vals[j], x);
      }
      if (MEAN && j1 > j0)
        acc = make_float4(__fdiv_rn(acc.x, cnt), __fdiv_rn(acc.y, cnt), __fdiv_rn(acc.z, cnt), __fdiv_rn(acc.w, cnt));
      reinterpret_cast<float4*>(dst + (long long)r * ld_dst)[i] = acc;
    }
  } else {
    for (long long i = (long long)blockIdx.x * MIX_THREADS + threadIdx.x; i < n;
         i += (long long)gridDim.x * MIX_THREADS) {
      float acc = 0.f;
      for (int j = j0; j < j1; ++j) {
        const float x = src[(long long)col[j] * ld_src + i];
        const float t = MEAN ? x : __fmul_rn(x, vals[j]);
        acc = (j == j0) ? t : __fadd_rn(acc, t);
      }
      if (MEAN && j1 > j0) acc = __fdiv_rn(acc, cnt);
      dst[(long long)r * ld_dst + i] = acc;
    }
  }
}

// out[:] = sum_c term(c), c in ascending order, term(c) = x_c * w[c] (MODE 0: Cloud.run, fedavg_aggregate),
// x_c / div (MODE 1: `p[key] += paras[key] / len(client_list)`, FLGAN/MNIST/flgan.py:151-158) or x_c with the sum divided
// by div at the end (MODE 2: receive_parameter, CGLGAN/2DMG/main.py:171-179). Every product / quotient and every sum is
// rounded separately, in the order the reference's dict loop applies them, so the result is order-exact. The sum over c
// is a serial fp32 chain per element: the column index is the only parallel axis, and the vector width VW (floats per
// thread) is chosen by the host so that a short row (a generator trunk: 179 k floats) still spreads over > 100 k threads
// with U independent loads in flight each.
template <int VW> struct VecT;
template <> struct VecT<4> { typedef float4 type; };
template <> struct VecT<2> { typedef float2 type; };
template <> struct VecT<1> { typedef float type; };
template <int VW>
__device__ __forceinline__ void ld_vec(const float* p, float (&v)[VW]) {
  if (VW == 4) {
    const float4 t = ld_stream(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2 % VW] = t.z; v[3 % VW] = t.w;
  } else if (VW == 2) {
    float2 t;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(t.x), "=f"(t.y) : "l"(p));
    v[0] = t.x; v[1 % VW] = t.y;
  } else {
    float t;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(t) : "l"(p));
    v[0] = t;
  }
}
template <int MODE>
__device__ __forceinline__ float wsum_term(float x, float w, float div) {
  return MODE == 0 ? __fmul_rn(x, w) : (MODE == 1 ? __fdiv_rn(x, div) : x);
}
template <int VW, int MODE>
__global__ void __launch_bounds__(MIX_THREADS) wsum_kernel(int C, long long n, const float* __restrict__ w, float div,
                                                          const int* __restrict__ rows,
                                                          const float* __restrict__ src, long long ld_src,
                                                          float* __restrict__ out) {
  constexpr int U = (VW == 4) ? 8 : 16;
  const long long nv = n / VW;
  for (long long i = (long long)blockIdx.x * MIX_THREADS + threadIdx.x; i < nv; i += (long long)gridDim.x * MIX_THREADS) {
    float acc[VW];
#pragma unroll
    for (int e = 0; e < VW; ++e) acc[e] = 0.f;
    int c = 0;
    if (C > 0) {
      float x[VW];
      ld_vec<VW>(src + (long long)(rows ? rows[0] : 0) * ld_src + i * VW, x);
      const float ww = MODE == 0 ? w[0] : 0.f;
#pragma unroll
      for (int e = 0; e < VW; ++e) acc[e] = wsum_term<MODE>(x[e], ww, div);   // the first term initialises (p[key] = ...)
      c = 1;
    }
    for (; c + U - 1 < C; c += U) {
      float x[U][VW];
#pragma unroll
      for (int u = 0; u < U; ++u) ld_vec<VW>(src + (long long)(rows ? rows[c + u] : c + u) * ld_src + i * VW, x[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float ww = MODE == 0 ? w[c + u] : 0.f;
#pragma unroll
        for (int e = 0; e < VW; ++e) acc[e] = __fadd_rn(acc[e], wsum_term<MODE>(x[u][e], ww, div));
      }
    }
    for (; c < C; ++c) {
      float x[VW];
      ld_vec<VW>(src + (long long)(rows ? rows[c] : c) * ld_src + i * VW, x);
      const float ww = MODE == 0 ? w[c] : 0.f;
#pragma unroll
      for (int e = 0; e < VW; ++e) acc[e] = __fadd_rn(acc[e], wsum_term<MODE>(x[e], ww, div));
    }
#pragma unroll
    for (int e = 0; e < VW; ++e) out[i * VW + e] = (MODE == 2) ? __fdiv_rn(acc[e], div) : acc[e];
  }
}

// dst[row,:] = sigma*dst[row,:] + (1-sigma)*g[:]   (recv_p = segema*self_p + (1-segema)*recv_p)
template <bool VEC>
__global__ void __launch_bounds__(MIX_THREADS) bcast_mix_kernel(long long n, const int* __restrict__ rows, float sigma,
                                                               float one_minus_sigma, const float* __restrict__ gsrc,
                                                               float* __restrict__ dst, long long ld_dst) {
  const int r = blockIdx.y;
  float* d = dst + (long long)(rows ? rows[r] : r) * ld_dst;
  if (VEC) {
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * MIX_THREADS + threadIdx.x; i < n4;
         i += (long long)gridDim.x * MIX_THREADS) {
      float4 gv = reinterpret_cast<const float4*>(gsrc)[i];
      float4 o;
      if (sigma == 0.f) {
        // segema == 0: 0*self + 1*recv == recv exactly (for finite self); skip the read
        o = gv;
      } else {
        float4 s = reinterpret_cast<float4*>(d)[i];
        o.x = __fadd_rn(__fmul_rn(sigma, s.x), __fmul_rn(one_minus_sigma, gv.x));
        o.y = __fadd_rn(__fmul_rn(sigma, s.y), __fmul_rn(one_minus_sigma, gv.y));
        o.z = __fadd_rn(__fmul_rn(sigma, s.z), __fmul_rn(one_minus_sigma, gv.z));
        o.w = __fadd_rn(__fmul_rn(sigma, s.w), __fmul_rn(one_minus_sigma, gv.w));
      }
      reinterpret_cast<float4*>(d)[i] = o;
    }
  } else {
    for (long long i = (long long)blockIdx.x * MIX_THREADS + threadIdx.x; i < n;
         i += (long long)gridDim.x * MIX_THREADS) {
      float gv = gsrc[i];
      d[i] = (sigma == 0.f) ? gv : __fadd_rn(__fmul_rn(sigma, d[i]), __fmul_rn(one_minus_sigma, gv));
    }
  }
}

// Adam over R packed rows with an explicit gradient buffer.
template <bool VEC>
__global__ void __launch_bounds__(MIX_THREADS) adam_rows_kernel(long long n, long long ld, float* __restrict__ p,
                                                               const float* __restrict__ g, float* __restrict__ m,
                                                               float* __restrict__ v, int* __restrict__ step, float lr,
                                                               float b1, float b2, float eps) {
  const int r = blockIdx.y;
  // every CTA of the row derives the same t; the row's first CTA publishes it after a grid-wide
  // agreement is unnecessary because step[] is only written by the follow-up bump kernel.
  const AdamScalars s = make_adam_scalars(step[r] + 1, lr, b1, b2, eps);
  const long long base = (long long)r * ld;
  if (VEC) {
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * MIX_THREADS + threadIdx.x; i < n4;
         i += (long long)gridDim.x * MIX_THREADS) {
      float4 p4 = reinterpret_cast<float4*>(p + base)[i];
      float4 g4 = reinterpret_cast<const float4*>(g + base)[i];
      float4 m4 = reinterpret_cast<float4*>(m + base)[i];
      float4 v4 = reinterpret_cast<float4*>(v + base)[i];
      adam_update(p4.x, m4.x, v4.x, g4.x, s);
      adam_update(p4.y, m4.y, v4.y, g4.y, s);
      adam_update(p4.z, m4.z, v4.z, g4.z, s);
      adam_update(p4.w, m4.w, v4.w, g4.w, s);
      reinterpret_cast<float4*>(p + base)[i] = p4;
      reinterpret_cast<float4*>(m + base)[i] = m4;
      reinterpret_cast<float4*>(v + base)[i] = v4;
    }
  } else {
    for (long long i = (long long)blockIdx.x * MIX_THREADS + threadIdx.x; i < n;
         i += (long long)gridDim.x * MIX_THREADS) {
      float pp = p[base + i], mm = m[base + i], vv = v[base + i];
      adam_update(pp, mm, vv, g[base + i], s);
      p[base + i] = pp; m[base + i] = mm; v[base + i] = vv;
    }
  }
}
__global__ void bump_rows_kernel(int R, int* step) {
  int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r < R) step[r] += 1;
}

// out[s,:] = sum_j weights[c_j] * dxg[c_j,:], c_j = clients[j], j in [srv_ptr[s], srv_ptr[s+1])
template <bool VEC>
__global__ void __launch_bounds__(MIX_THREADS) dxg_reduce_kernel(long long n, const int* __restrict__ srv_ptr,
                                                                const int* __restrict__ clients,
                                                                const float* __restrict__ weights,
                                                                const float* __restrict__ dxg,
                                                                float* __restrict__ out) {
  const int s = blockIdx.y;
  const int j0 = srv_ptr[s], j1 = srv_ptr[s + 1];
  if (VEC) {
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * MIX_THREADS + threadIdx.x; i < n4;
         i += (long long)gridDim.x * MIX_THREADS) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int j = j0; j < j1; ++j) {
        int c = clients ? clients[j] : j;
        float w = weights ? weights[c] : 1.f;
        float4 x = ld_stream(reinterpret_cast<const float4*>(dxg + (long long)c * n) + i);
        acc.x = fmaf(w, x.x, acc.x); acc.y = fmaf(w, x.y, acc.y);
        acc.z = fmaf(w, x.z, acc.z); acc.w = fmaf(w, x.w, acc.w);
      }
      reinterpret_cast<float4*>(out + (long long)s * n)[i] = acc;
    }
  } else {
    for (long long i = (long long)blockIdx.x * MIX_THREADS + threadIdx.x; i < n;
         i += (long long)gridDim.x * MIX_THREADS) {
      float acc = 0.f;
      for (int j = j0; j < j1; ++j) {
        int c = clients ? clients[j] : j;
        float w = weights ? weights[c] : 1.f;
        acc = fmaf(w, dxg[(long long)c * n + i], acc);
      }
      out[(long long)s * n + i] = acc;
    }
  }
}

static inline int grid_x(long long work_items, int rows) {
  // enough CTAs for ~8 resident CTAs on each of the 148 SMs, split over the `rows` axis
  long long want = (work_items + MIX_THREADS - 1) / MIX_THREADS;
  long long cap = (148LL * 8 + rows - 1) / rows;
  if (cap < 1) cap = 1;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

}  // namespace cgl

using namespace cgl;

extern "C" int cgl_mix_csr(int R, int64_t n, const int32_t* row_ptr, const int32_t* col, const float* vals,
                           const float* src, int64_t ld_src, float* dst, int64_t ld_dst, cgl_stream_t stream) {
  if (R == 0 || n == 0) return CGL_OK;
  CGL_REQUIRE(R > 0 && R <= 65535 && n > 0, "bad shape R=%d n=%lld", R, (long long)n);
  CGL_REQUIRE(row_ptr && col && src && dst, "NULL tensor pointer");   // vals == NULL: row means
  CGL_REQUIRE(src != dst, "cgl_mix_csr cannot run in place");
  cudaStream_t st = (cudaStream_t)stream;
  bool vec = aligned16(src) && aligned16(dst) && n % 4 == 0 && ld_src % 4 == 0 && ld_dst % 4 == 0;
  dim3 grid(grid_x(vec ? n / 4 : n, R), R);
  // the column count lives on the device: at least one source row is read per written row
  ProfScope prof(CGL_PROF_MIX, 8.0 * R * (double)n, 0.0, st);
  if (vals) {
    if (vec) mix_csr_kernel<true, false><<<grid, MIX_THREADS, 0, st>>>(n, row_ptr, col, vals, src, ld_src, dst, ld_dst);
    else mix_csr_kernel<false, false><<<grid, MIX_THREADS, 0, st>>>(n, row_ptr, col, vals, src, ld_src, dst, ld_dst);
  } else {
    if (vec) mix_csr_kernel<true, true><<<grid, MIX_THREADS, 0, st>>>(n, row_ptr, col, vals, src, ld_src, dst, ld_dst);
    else mix_csr_kernel<false, true><<<grid, MIX_THREADS, 0, st>>>(n, row_ptr, col, vals, src, ld_src, dst, ld_dst);
  }
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

template <int MODE>
static int launch_wsum(int C, int64_t n, const float* w, float div, const int32_t* rows, const float* src, int64_t ld_src,
                       float* out, cudaStream_t st) {
  // widest vector the pointers allow, narrowed until the row spreads over >= ~100 k threads (or VW reaches 1)
  int vw = (aligned16(src) && aligned16(out) && n % 4 == 0 && ld_src % 4 == 0) ? 4
           : ((((uintptr_t)src | (uintptr_t)out) & 7u) == 0 && n % 2 == 0 && ld_src % 2 == 0) ? 2 : 1;
  while (vw > 1 && n / vw < 100000) vw >>= 1;
  const long long items = n / vw;
  const int gx = (int)((items + MIX_THREADS - 1) / MIX_THREADS);
  ProfScope prof(CGL_PROF_MIX, 4.0 * ((double)C + 1.0) * (double)n, 2.0 * C * (double)n, st);   // 4 P (C_in + R_out)
  if (vw == 4) wsum_kernel<4, MODE><<<gx, MIX_THREADS, 0, st>>>(C, n, w, div, rows, src, ld_src, out);
  else if (vw == 2) wsum_kernel<2, MODE><<<gx, MIX_THREADS, 0, st>>>(C, n, w, div, rows, src, ld_src, out);
  else wsum_kernel<1, MODE><<<gx, MIX_THREADS, 0, st>>>(C, n, w, div, rows, src, ld_src, out);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

extern "C" int cgl_wsum(int C, int64_t n, const float* w, const int32_t* rows, const float* src, int64_t ld_src,
                        float* out, cgl_stream_t stream) {
  if (n == 0) return CGL_OK;
  CGL_REQUIRE(C >= 0 && n > 0, "bad shape C=%d n=%lld", C, (long long)n);
  CGL_REQUIRE((C == 0 || (w && src)) && out, "NULL tensor pointer");
  return launch_wsum<0>(C, n, w, 1.f, rows, src, ld_src, out, (cudaStream_t)stream);
}

extern "C" int cgl_wsum_div(int C, int64_t n, float divisor, int sum_first, const int32_t* rows, const float* src,
                            int64_t ld_src, float* out, cgl_stream_t stream) {
  if (n == 0) return CGL_OK;
  CGL_REQUIRE(C >= 0 && n > 0, "bad shape C=%d n=%lld", C, (long long)n);
  CGL_REQUIRE((C == 0 || src) && out, "NULL tensor pointer");
  CGL_REQUIRE(divisor != 0.f, "divisor is zero");
  if (sum_first) return launch_wsum<2>(C, n, nullptr, divisor, rows, src, ld_src, out, (cudaStream_t)stream);
  return launch_wsum<1>(C, n, nullptr, divisor, rows, src, ld_src, out, (cudaStream_t)stream);
}

extern "C" int cgl_bcast_mix(int R, int64_t n, const int32_t* rows, float sigma, const float* g, float* dst,
                             int64_t ld_dst, cgl_stream_t stream) {
  if (R == 0 || n == 0) return CGL_OK;
  CGL_REQUIRE(R > 0 && R <= 65535 && n > 0, "bad shape R=%d n=%lld", R, (long long)n);
  CGL_REQUIRE(g && dst, "NULL tensor pointer");
  cudaStream_t st = (cudaStream_t)stream;
  bool vec = aligned16(g) && aligned16(dst) && n % 4 == 0 && ld_dst % 4 == 0;
  dim3 grid(grid_x(vec ? n / 4 : n, R), R);
  // torch evaluates (1 - segema) in the tensor dtype (fp32) when segema is a tensor, in double
  // when it is a python float; both round to the same fp32 for the reference's 0 / 0.5 / 1.
  float oms = (float)(1.0 - (double)sigma);
  ProfScope prof(CGL_PROF_MIX, 4.0 * (double)n * ((sigma != 0.f ? 2.0 : 1.0) * R + 1.0), 0.0, st);
  if (vec) bcast_mix_kernel<true><<<grid, MIX_THREADS, 0, st>>>(n, rows, sigma, oms, g, dst, ld_dst);
  else bcast_mix_kernel<false><<<grid, MIX_THREADS, 0, st>>>(n, rows, sigma, oms, g, dst, ld_dst);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

extern "C" int cgl_adam_rows(int R, int64_t n, int64_t ld, float* p, const float* g, float* m, float* v,
                             int32_t* step, float lr, float beta1, float beta2, float eps, cgl_stream_t stream) {
  if (R == 0 || n == 0) return CGL_OK;
  CGL_REQUIRE(R > 0 && R <= 65535 && n > 0 && ld >= n, "bad shape R=%d n=%lld ld=%lld", R, (long long)n, (long long)ld);
  CGL_REQUIRE(p && g && m && v && step, "NULL tensor pointer");
  cudaStream_t st = (cudaStream_t)stream;
  bool vec = aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v) && n % 4 == 0 && ld % 4 == 0;
  dim3 grid(grid_x(vec ? n / 4 : n, R), R);
  if (vec) adam_rows_kernel<true><<<grid, MIX_THREADS, 0, st>>>(n, ld, p, g, m, v, step, lr, beta1, beta2, eps);
  else adam_rows_kernel<false><<<grid, MIX_THREADS, 0, st>>>(n, ld, p, g, m, v, step, lr, beta1, beta2, eps);
  CGL_CHECK_LAUNCH();
  bump_rows_kernel<<<(R + 127) / 128, 128, 0, st>>>(R, step);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

extern "C" int cgl_dxg_reduce(int S, const int32_t* srv_ptr, const int32_t* clients, const float* weights,
                              const float* dxg, int64_t n, float* out, cgl_stream_t stream) {
  if (S == 0 || n == 0) return CGL_OK;
  CGL_REQUIRE(S > 0 && S <= 65535 && n > 0, "bad shape S=%d n=%lld", S, (long long)n);
  CGL_REQUIRE(srv_ptr && dxg && out, "NULL tensor pointer");
  cudaStream_t st = (cudaStream_t)stream;
  bool vec = aligned16(dxg) && aligned16(out) && n % 4 == 0;
  dim3 grid(grid_x(vec ? n / 4 : n, S), S);
  ProfScope prof(CGL_PROF_ELEMENTWISE, 0.0, 0.0, st);
  if (vec) dxg_reduce_kernel<true><<<grid, MIX_THREADS, 0, st>>>(n, srv_ptr, clients, weights, dxg, out);
  else dxg_reduce_kernel<false><<<grid, MIX_THREADS, 0, st>>>(n, srv_ptr, clients, weights, dxg, out);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}
