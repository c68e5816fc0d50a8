// gemm.cuh -- grouped fp32 GEMM over G independent groups (one group = one simulated client /
// edge server) with the epilogues the GAN step needs. FFMA path (exact fp32 semantics):
//   C[g][m][n] = sum_k Aop[g][m][k] * Bop[g][n][k]
// 128x128x16 CTA tile, 256 threads, 8x8 register micro-tile, double-buffered shared memory with
// register prefetch of the next k-tile. Operands are addressed through a RowMap so that one
// operand can be the concatenation of two row blocks (real batch | fake batch, reference
// CGLGAN/2DMG/main.py:361-363) and so that packed parameter rows can be picked through an index.
#pragma once
#include "common.cuh"

namespace cgl {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 16;
constexpr int GEMM_THREADS = 256;
constexpr int LDS = 132;  // padded leading dimension of the [BK][128] shared tiles

// (group g, row r) -> pointer to a row of `ld` floats.
//   r <  rows0 : base0 + (idx0 ? idx0[g] : g) * gstride0 + r * ld
//   r >= rows0 : base1 + (idx1 ? idx1[g] : g) * gstride1 + (r - rows0) * ld
struct RowMap {
  const float* base0;
  long long gstride0;
  const int* idx0;
  int rows0;
  const float* base1;
  long long gstride1;
  const int* idx1;
  int ld;
  int vec;  // host-verified: every row start is 16B aligned and ld % 4 == 0
};

static inline RowMap single_rows(const float* base, long long gstride, const int* idx, int ld) {
  RowMap m;
  m.base0 = base; m.gstride0 = gstride; m.idx0 = idx; m.rows0 = 0x7fffffff;
  m.base1 = base; m.gstride1 = 0; m.idx1 = nullptr;
  m.ld = ld;
  m.vec = (aligned16(base) && (gstride % 4 == 0) && (ld % 4 == 0)) ? 1 : 0;
  return m;
}
static inline RowMap dual_rows(const float* b0, long long gs0, const int* i0, int rows0,
                               const float* b1, long long gs1, const int* i1, int ld) {
  RowMap m;
  m.base0 = b0; m.gstride0 = gs0; m.idx0 = i0; m.rows0 = rows0;
  m.base1 = b1; m.gstride1 = gs1; m.idx1 = i1;
  m.ld = ld;
  m.vec = (aligned16(b0) && aligned16(b1) && (gs0 % 4 == 0) && (gs1 % 4 == 0) && (ld % 4 == 0)) ? 1 : 0;
  return m;
}

struct Rows {  // a RowMap resolved for one group
  const float* p0;
  const float* p1;  // pre-offset by -rows0*ld so that p1 + r*ld is right for r >= rows0
  int rows0;
  int ld;
};
__device__ __forceinline__ Rows resolve(const RowMap& m, int g) {
  Rows r;
  int i0 = m.idx0 ? m.idx0[g] : g;
  r.p0 = m.base0 + (long long)i0 * m.gstride0;
  r.rows0 = m.rows0;
  r.ld = m.ld;
  if (m.rows0 != 0x7fffffff) {
    int i1 = m.idx1 ? m.idx1[g] : g;
    r.p1 = m.base1 + (long long)i1 * m.gstride1 - (long long)m.rows0 * m.ld;
  } else {
    r.p1 = r.p0;
  }
  return r;
}
__device__ __forceinline__ const float* row_ptr(const Rows& r, int row) {
  return (row < r.rows0 ? r.p0 : r.p1) + (long long)row * r.ld;
}

// ---- tile loaders: global -> 8 registers per thread -> shared [BK][LDS] --------------------
// KMAJOR: the tile dimension indexes rows, the contraction index runs along a row (contiguous).
// !KMAJOR: the contraction index picks the row, the tile dimension runs along it (contiguous).
template <bool KMAJOR>
__device__ __forceinline__ void tile_load(const Rows& R, bool vec, int t0, int dimT, int k0, int dimK,
                                          float (&reg)[8]) {
  const int t = threadIdx.x;
  if (KMAJOR) {
    if (vec) {
      const int q = t & 3, r = t >> 2;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        int row = t0 + r + 64 * p;
        int k = k0 + q * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < dimT && k < dimK) v = __ldg(reinterpret_cast<const float4*>(row_ptr(R, row) + k));
        reg[4 * p + 0] = v.x; reg[4 * p + 1] = v.y; reg[4 * p + 2] = v.z; reg[4 * p + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        int e = p * GEMM_THREADS + t;
        int kk = e & 15, r = e >> 4;
        int row = t0 + r, k = k0 + kk;
        reg[p] = (row < dimT && k < dimK) ? __ldg(row_ptr(R, row) + k) : 0.f;
      }
    }
  } else {
    if (vec) {
      const int c4 = t & 31, kk = t >> 5;
#pragma unroll
      for (int p = 0; p < 2; ++p) {
        int k = k0 + kk + 8 * p;
        int col = t0 + c4 * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < dimK && col < dimT) v = __ldg(reinterpret_cast<const float4*>(row_ptr(R, k) + col));
        reg[4 * p + 0] = v.x; reg[4 * p + 1] = v.y; reg[4 * p + 2] = v.z; reg[4 * p + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        int e = p * GEMM_THREADS + t;
        int c = e & 127, kk = e >> 7;
        int k = k0 + kk, col = t0 + c;
        reg[p] = (k < dimK && col < dimT) ? __ldg(row_ptr(R, k) + col) : 0.f;
      }
    }
  }
}

template <bool KMAJOR>
__device__ __forceinline__ void tile_store(float (*S)[LDS], bool vec, const float (&reg)[8]) {
  const int t = threadIdx.x;
  if (KMAJOR) {
    if (vec) {
      const int q = t & 3, r = t >> 2;
#pragma unroll
      for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int j = 0; j < 4; ++j) S[q * 4 + j][r + 64 * p] = reg[4 * p + j];
    } else {
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        int e = p * GEMM_THREADS + t;
        S[e & 15][e >> 4] = reg[p];
      }
    }
  } else {
    if (vec) {
      const int c4 = t & 31, kk = t >> 5;
#pragma unroll
      for (int p = 0; p < 2; ++p)
        *reinterpret_cast<float4*>(&S[kk + 8 * p][c4 * 4]) =
            make_float4(reg[4 * p + 0], reg[4 * p + 1], reg[4 * p + 2], reg[4 * p + 3]);
    } else {
#pragma unroll
      for (int p = 0; p < 8; ++p) {
        int e = p * GEMM_THREADS + t;
        S[e >> 7][e & 127] = reg[p];
      }
    }
  }
}

// ---- epilogues -----------------------------------------------------------------------------
enum { EPI_FWD = 0, EPI_BWD_DATA = 1, EPI_ADAM = 2, EPI_STORE = 3 };

struct GemmParams {
  int M, N, K;  // C is [M,N]; K is the contraction length
  RowMap A, B;
  // output C[g][m][n] at cbase + (cidx ? cidx[g] : g) * c_gstride + c_off + m*ldc + n
  float* cbase;
  long long c_gstride;
  const int* cidx;
  long long c_off;
  int ldc;
  int c_vec;
  // EPI_FWD: bias[n] at bias_base + row(g) * bias_gstride + bias_off; activation
  const float* bias_base;
  long long bias_gstride;
  const int* bias_idx;
  long long bias_off;
  int act;
  float slope;
  // EPI_BWD_DATA: saved activations, same indexing as C but own base/stride (NULL: no derivative)
  const float* saved;
  long long saved_gstride;
  // EPI_ADAM: C is the weight matrix inside the packed row; m/v share its indexing. The CTAs of
  // n-tile 0 also reduce the bias gradient (column sums of the A operand) and update the bias.
  float* adam_m;
  float* adam_v;
  const int* step;  // step[row(g)], already incremented for this update
  const AdamScalars* scal;  // [G] precomputed scalars of this step (NULL: derive from step)
  float lr, b1, b2, eps;
  // EPI_STORE with bias gradient: db[m] = sum_k Aop[m][k] stored at cbase + ... + dbias_off (or -1)
  long long dbias_off;
};

template <bool A_KMAJOR, bool B_KMAJOR, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS) grouped_gemm_kernel(const GemmParams p) {
  __shared__ __align__(16) float As[2][BK][LDS];
  __shared__ __align__(16) float Bs[2][BK][LDS];

  const int g = blockIdx.z;
  const int m0 = blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  const Rows RA = resolve(p.A, g);
  const Rows RB = resolve(p.B, g);
  // vector loads need 4-float granularity along the contiguous direction
  const bool vecA = p.A.vec && (A_KMAJOR ? (p.K % 4 == 0) : (p.M % 4 == 0));
  const bool vecB = p.B.vec && (B_KMAJOR ? (p.K % 4 == 0) : (p.N % 4 == 0));

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float ra[8], rb[8];
  const bool want_bias = (EPI == EPI_ADAM || EPI == EPI_STORE) && blockIdx.x == 0 &&
                         (EPI == EPI_ADAM ? p.bias_off >= 0 : p.dbias_off >= 0);
  float bsum = 0.f;

  const int nk = (p.K + BK - 1) / BK;
  tile_load<A_KMAJOR>(RA, vecA, m0, p.M, 0, p.K, ra);
  tile_load<B_KMAJOR>(RB, vecB, n0, p.N, 0, p.K, rb);
  tile_store<A_KMAJOR>(As[0], vecA, ra);
  tile_store<B_KMAJOR>(Bs[0], vecB, rb);
  __syncthreads();

  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) {
      tile_load<A_KMAJOR>(RA, vecA, m0, p.M, (kt + 1) * BK, p.K, ra);
      tile_load<B_KMAJOR>(RB, vecB, n0, p.N, (kt + 1) * BK, p.K, rb);
    }
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (want_bias && threadIdx.x < BM) {
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) bsum += As[cur][kk][threadIdx.x];
    }
    if (kt + 1 < nk) {
      tile_store<A_KMAJOR>(As[cur ^ 1], vecA, ra);
      tile_store<B_KMAJOR>(Bs[cur ^ 1], vecB, rb);
    }
    __syncthreads();
  }

  // ---- epilogue ----
  const int rowid = p.cidx ? p.cidx[g] : g;
  float* C = p.cbase + (long long)rowid * p.c_gstride + p.c_off;

  if (EPI == EPI_FWD) {
    const int brow = p.bias_idx ? p.bias_idx[g] : g;
    const float* bias = p.bias_base ? p.bias_base + (long long)brow * p.bias_gstride + p.bias_off : nullptr;
    float bv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      bv[j] = (bias && n < p.N) ? __ldg(bias + n) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      if (m >= p.M) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int n = n0 + h * 64 + tx * 4;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) o[j] = act_fwd(acc[i][h * 4 + j] + bv[h * 4 + j], p.act, p.slope);
        float* dst = C + (long long)m * p.ldc + n;
        if (p.c_vec && n + 3 < p.N) {
          *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < p.N) dst[j] = o[j];
        }
      }
    }
  } else if (EPI == EPI_BWD_DATA || EPI == EPI_STORE) {
    const float* S = (EPI == EPI_BWD_DATA && p.saved) ? p.saved + (long long)g * p.saved_gstride : nullptr;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      if (m >= p.M) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int n = n0 + h * 64 + tx * 4;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v = acc[i][h * 4 + j];
          if (S && n + j < p.N) v *= act_bwd_from_out(__ldg(S + (long long)m * p.ldc + n + j), p.act, p.slope);
          o[j] = v;
        }
        float* dst = C + (long long)m * p.ldc + n;
        if (p.c_vec && n + 3 < p.N) {
          *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < p.N) dst[j] = o[j];
        }
      }
    }
    if (EPI == EPI_STORE && want_bias && threadIdx.x < BM) {
      int m = m0 + threadIdx.x;
      if (m < p.M) (p.cbase + (long long)rowid * p.c_gstride + p.dbias_off)[m] = bsum;
    }
  } else {  // EPI_ADAM
    const AdamScalars s = p.scal ? p.scal[g] : make_adam_scalars(p.step[rowid], p.lr, p.b1, p.b2, p.eps);
    float* Mo = p.adam_m + (long long)rowid * p.c_gstride + p.c_off;
    float* Vo = p.adam_v + (long long)rowid * p.c_gstride + p.c_off;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      if (m >= p.M) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        int n = n0 + h * 64 + tx * 4;
        long long off = (long long)m * p.ldc + n;
        if (p.c_vec && n + 3 < p.N) {
          float4 w4 = *reinterpret_cast<float4*>(C + off);
          float4 m4 = *reinterpret_cast<float4*>(Mo + off);
          float4 v4 = *reinterpret_cast<float4*>(Vo + off);
          adam_update(w4.x, m4.x, v4.x, acc[i][h * 4 + 0], s);
          adam_update(w4.y, m4.y, v4.y, acc[i][h * 4 + 1], s);
          adam_update(w4.z, m4.z, v4.z, acc[i][h * 4 + 2], s);
          adam_update(w4.w, m4.w, v4.w, acc[i][h * 4 + 3], s);
          *reinterpret_cast<float4*>(C + off) = w4;
          *reinterpret_cast<float4*>(Mo + off) = m4;
          *reinterpret_cast<float4*>(Vo + off) = v4;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n + j < p.N) {
              float w = C[off + j], mm = Mo[off + j], vv = Vo[off + j];
              adam_update(w, mm, vv, acc[i][h * 4 + j], s);
              C[off + j] = w; Mo[off + j] = mm; Vo[off + j] = vv;
            }
          }
        }
      }
    }
    if (want_bias && threadIdx.x < BM) {
      int m = m0 + threadIdx.x;
      if (m < p.M) {
        long long boff = (long long)rowid * p.c_gstride + p.bias_off + m;
        float w = p.cbase[boff], mm = p.adam_m[boff], vv = p.adam_v[boff];
        adam_update(w, mm, vv, bsum, s);
        p.cbase[boff] = w; p.adam_m[boff] = mm; p.adam_v[boff] = vv;
      }
    }
  }
}

template <bool A_KMAJOR, bool B_KMAJOR, int EPI>
static inline cudaError_t launch_grouped_gemm(const GemmParams& p, int G, cudaStream_t stream) {
  if (G <= 0 || p.M <= 0 || p.N <= 0) return cudaSuccess;
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, G);
  grouped_gemm_kernel<A_KMAJOR, B_KMAJOR, EPI><<<grid, GEMM_THREADS, 0, stream>>>(p);
  count_launch();
  return cudaGetLastError();
}

}  // namespace cgl
