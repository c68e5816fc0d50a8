// client_fused.cuh -- K1: the whole client step of a SMALL discriminator in ONE launch, one CTA per client, with the
// client's weights resident in shared memory (north-star part 1; Worker.train, CGLGAN/2DMG/main.py:344-375 with
// D = CGLGAN/2DMG/model.py:54-71, 2 -> 128 -> 256 -> 1, 33,665 parameters = 135 KB).
//
//   D step  : X = [real | fake] (2B rows) -> h1 -> h2 -> logit -> sigmoid -> BCE (or MSE), backward through all three
//             layers, torch.optim.Adam on every parameter -- W / m / v cross HBM once (24 B per parameter), the weights
//             are read from HBM once per launch and then live in shared memory for both parts
//   G loss  : with the UPDATED weights, loss(D(Xg), valid) and dLoss/dXg (the hand-off to the server, a3)
//
// Rows are processed in chunks of K1_R = 48 (activations of a chunk live in shared memory: 76 KB next to the 135 KB of
// weights). The three products with the 128 x 256 layer (98 % of the MACs) run on the tensor cores through the
// WARP-LEVEL path, mma.sync.m16n8k8 tf32 with the 3xTF32 split done in registers on the fragments (hi = tf32(x) rounded,
// lo = x - hi exact; lo*hi + hi*lo + hi*hi accumulated in fp32 registers):
//   * the operands are read straight from the resident fp32 copies with plain LDS -- no descriptor layouts. tcgen05
//     cannot keep W2 resident: its hi / lo images are 256 KB, and the forward product needs W2 K-major while the data
//     gradient needs it MN-major (two different shared-memory images, tc_gemm.cuh), so a tcgen05 version would restage
//     W2 for every chunk like the layered kernels do;
//   * the weight gradient of the 128 x 256 layer is accumulated over the chunks in REGISTERS (64 per thread).
// Measured on B200 (profiles/experiments/mma_sync_probe.cu): mma.sync tf32 278 TFLOP/s = 93 TFLOP/s for 3xTF32 products,
// FFMA 72 TFLOP/s; the first versions of this kernel with FFMA phases reached 27 TFLOP/s (profiles/ncu_k1_ffma_r2.md).
//
// Phases per chunk (512 threads = 16 warps; gid = lane / 4, tig = lane % 4 of the mma fragments):
//   A  h1 = lrelu(x W1^T + b1)                      FFMA; thread: column i, every fourth row
//   B  h2 = lrelu(h1 W2^T + b2)                     mma: M = rows (3 tiles), N = out (warp: 16 columns), K = 128
//   C  logit, loss term, dlogit; dW3 / db3 / db2;   warp per row, then thread per column; h2 is overwritten by dZ2
//   E  dZ1 = (dZ2 W2) * lrelu'(h1)                  mma: M = rows, N = in (warp: 16 columns), K = HALF of 256 (the consumers are
//                                                   linear in dZ1: the halves meet there), the k index of a step permuted
//                                                   (slot t <-> k 2t, 2t+1) so that both fragment loads are conflict-free;
//                                                   feeds dW1 / db1 (D step) or dXg (G loss) from registers
//   D  dW2 += dZ2^T h1                              mma: M = out (warp: 64), N = in (warp: 32), K = rows of the chunk
// Shared-memory row strides: W2 and h1 132 floats, h2 / dZ2 264 floats (fragment loads hit 32 different banks).
#pragma once
#include "linear.cuh"

namespace cgl {

constexpr int K1_H1 = 128;
constexpr int K1_H2 = 256;
constexpr int K1_R = 48;        // rows per chunk: three m16 tiles
constexpr int K1_THREADS = 512;
constexpr int K1_LDW = 132;     // floats per row of the resident W2 copy
constexpr int K1_LD1 = 132;     // ... of h1
constexpr int K1_LD2 = 264;     // ... of h2 / dZ2
constexpr int K1_MAXD = 2;      // widest input
constexpr int K1_MAXROWS = 512; // 2B
#ifndef K1_DEFAULT_ON
#define K1_DEFAULT_ON 0      // B200, 1024 clients: 1.62 ms per step against 1.47 ms of the layered tcgen05 kernels
#endif

struct K1Params {
  int d, B;
  float* params; float* adam_m; float* adam_v; long long ldp;
  const int* ids; int* step;
  long long w_off[3], b_off[3];
  const float* real; const int* n_real; const float* fake; const int* fake_idx;   // D step
  const float* xg; const int* xg_idx;                                             // G loss
  float* out_dloss; float* out_gloss; float* out_dxg;
  int loss_kind, last_act; float slope, d_scale;
  float lr, b1, b2, eps;
  int do_d, do_g;
};

constexpr int K1_SRED = K1_R * 16 * K1_MAXD;   // dXg: one partial sum per warp
constexpr int K1_SMEM_FLOATS = K1_H2 * K1_LDW + K1_R * K1_LD1 + K1_R * K1_LD2 + K1_R * K1_MAXD + K1_H1 * K1_MAXD + K1_H1 +
                               K1_H2 + K1_H2 + K1_R + K1_MAXROWS + K1_SRED + 16;
constexpr size_t K1_SMEM_BYTES = (size_t)K1_SMEM_FLOATS * sizeof(float);
static_assert(K1_SMEM_BYTES <= 227 * 1024, "K1 shared memory");

__device__ __forceinline__ float k1_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// D (16x8, fp32) += A (16x8, tf32, row) * B (8x8, tf32, col)
__device__ __forceinline__ void k1_mma(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// hi = x rounded to tf32 (as tf32_hi of tc_gemm.cuh), lo = x - hi (exact in fp32; the tensor core drops its low bits)
__device__ __forceinline__ void k1_split(float x, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
template <int N>
__device__ __forceinline__ void k1_split_n(const float (&x)[N], uint32_t (&hi)[N], uint32_t (&lo)[N]) {
#pragma unroll
  for (int i = 0; i < N; ++i) k1_split(x[i], hi[i], lo[i]);
}
// the three products of the split (small ones first) for NT tiles that share the A fragment, product-major: an mma
// that adds into an accumulator is never issued right behind the previous one into the same accumulator (a warp issues
// in order: back-to-back dependent mma left the tensor pipe 35 % busy)
template <int NT>
__device__ __forceinline__ void k1_mma3(float (&c)[NT][4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                        const uint32_t (&bh)[NT][2], const uint32_t (&bl)[NT][2]) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) k1_mma(c[nt], al, bh[nt]);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) k1_mma(c[nt], ah, bl[nt]);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) k1_mma(c[nt], ah, bh[nt]);
}

__global__ void __launch_bounds__(K1_THREADS, 1) client_step_fused_kernel(const K1Params p) {
  extern __shared__ __align__(16) float k1s[];
  float* sW2 = k1s;                               // [256][132]
  float* sH1 = sW2 + K1_H2 * K1_LDW;              // [R][132]
  float* sH2 = sH1 + K1_R * K1_LD1;               // [R][264]   h2, then dZ2
  float* sX = sH2 + K1_R * K1_LD2;                // [R][2]
  float* sW1 = sX + K1_R * K1_MAXD;               // [128][d]
  float* sb1 = sW1 + K1_H1 * K1_MAXD;             // [128]
  float* sb2 = sb1 + K1_H1;                       // [256]
  float* sw3 = sb2 + K1_H2;                       // [256]
  float* sdz3 = sw3 + K1_H2;                      // [R]
  float* sloss = sdz3 + K1_R;                     // [2B]
  float* sred = sloss + K1_MAXROWS;               // [R][16 warps][d]: dXg partial sums
  float* smisc = sred + K1_SRED;                  // [0] = b3, [8..16) = AdamScalars of this step

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int gid = lane >> 2, tig = lane & 3;
  const int g = blockIdx.x;
  const int rowid = p.ids ? p.ids[g] : g;
  const int d = p.d, B = p.B;
  float* P = p.params + (long long)rowid * p.ldp;

  // ---- the client's weights: HBM -> shared, once ----
  {
    const float4* W2g = reinterpret_cast<const float4*>(P + p.w_off[1]);
    for (int i = tid; i < K1_H2 * (K1_H1 / 4); i += K1_THREADS) {
      const int o = i >> 5, k4 = i & 31;
      *reinterpret_cast<float4*>(sW2 + o * K1_LDW + 4 * k4) = W2g[i];
    }
    for (int i = tid; i < K1_H1 * d; i += K1_THREADS) sW1[i] = P[p.w_off[0] + i];
    if (tid < K1_H1) sb1[tid] = P[p.b_off[0] + tid];
    if (tid < K1_H2) {
      sb2[tid] = P[p.b_off[1] + tid];
      sw3[tid] = P[p.w_off[2] + tid];
    }
    if (tid == 0) smisc[0] = P[p.b_off[2]];
    // the step's Adam scalars (two double-precision pow): one thread, under the weight loads of the others
    if (tid == K1_THREADS - 1 && p.do_d)
      *reinterpret_cast<AdamScalars*>(smisc + 8) = make_adam_scalars(p.step[rowid] + 1, p.lr, p.b1, p.b2, p.eps);
  }
  __syncthreads();

  // ---- one chunk of rows through phases A, B, C (shared by both parts) ----
  // src rows: rr < rows0 -> p0 + rr * d, else p1 + (rr - rows0) * d
  auto phase_abc = [&](int c0, int nr, const float* p0, const float* p1, int rows0, int nv0, int n1, float t0, float t1,
                       float scale, bool train, float& gW3, float& gb2, float& gb3) {
    const int nm = (nr + 15) >> 4;   // m16 tiles that hold rows
    for (int i = tid; i < K1_R * d; i += K1_THREADS) {
      const int r = i / d, c = i - r * d;
      const int rr = c0 + r;
      float v = 0.f;
      if (r < nr) v = (rr < rows0) ? p0[(long long)rr * d + c] : p1[(long long)(rr - rows0) * d + c];
      sX[r * K1_MAXD + c] = v;
    }
    __syncthreads();
    {  // A
      const int i = tid & (K1_H1 - 1);
      float w[K1_MAXD];
#pragma unroll
      for (int c = 0; c < K1_MAXD; ++c) w[c] = (c < d) ? sW1[i * d + c] : 0.f;
      const float bb = sb1[i];
      for (int r = tid >> 7; r < K1_R; r += K1_THREADS / K1_H1) {
        float h = 0.f;
        if (r < nr) {
          float acc = 0.f;
#pragma unroll
          for (int c = 0; c < K1_MAXD; ++c)
            if (c < d) acc = fmaf(sX[r * K1_MAXD + c], w[c], acc);
          h = act_fwd(acc + bb, CGL_ACT_LRELU, p.slope);
        }
        sH1[r * K1_LD1 + i] = h;
      }
    }
    __syncthreads();
    {  // B: this warp's 16 output columns
      const int n0 = warp * 16;
      float acc[3][2][4];
#pragma unroll
      for (int mt = 0; mt < 3; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
      const float* wq = sW2 + (n0 + gid) * K1_LDW + tig;
      const float* hq = sH1 + gid * K1_LD1 + tig;
#pragma unroll 2
      for (int k0 = 0; k0 < K1_H1; k0 += 8) {
        uint32_t bh[2][2], bl[2][2];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const float bx[2] = {wq[nt * 8 * K1_LDW + k0], wq[nt * 8 * K1_LDW + k0 + 4]};
          k1_split_n(bx, bh[nt], bl[nt]);
        }
#pragma unroll
        for (int mt = 0; mt < 3; ++mt) {
          if (mt < nm) {
            const float* hr = hq + mt * 16 * K1_LD1 + k0;
            const float ax[4] = {hr[0], hr[8 * K1_LD1], hr[4], hr[8 * K1_LD1 + 4]};
            uint32_t ah[4], al[4];
            k1_split_n(ax, ah, al);
            k1_mma3<2>(acc[mt], ah, al, bh, bl);
          }
        }
      }
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) {
        if (mt < nm) {
#pragma unroll
          for (int nt = 0; nt < 2; ++nt) {
            const int o = n0 + nt * 8 + 2 * tig;
            const float b0 = sb2[o], b1 = sb2[o + 1];
            float* dst = sH2 + (mt * 16 + gid) * K1_LD2 + o;
            *reinterpret_cast<float2*>(dst) = make_float2(act_fwd(acc[mt][nt][0] + b0, CGL_ACT_LRELU, p.slope),
                                                          act_fwd(acc[mt][nt][1] + b1, CGL_ACT_LRELU, p.slope));
            *reinterpret_cast<float2*>(dst + 8 * K1_LD2) = make_float2(act_fwd(acc[mt][nt][2] + b0, CGL_ACT_LRELU, p.slope),
                                                                       act_fwd(acc[mt][nt][3] + b1, CGL_ACT_LRELU, p.slope));
          }
        }
      }
    }
    __syncthreads();
    // C: logits, loss terms, d loss / d logit (the arithmetic of head_kernel, dstep.cu)
    for (int r = warp; r < K1_R; r += K1_THREADS / 32) {
      if (r < nr) {
        float z = 0.f;
#pragma unroll
        for (int t = 0; t < K1_H2 / 32; ++t) z = fmaf(sH2[r * K1_LD2 + lane + 32 * t], sw3[lane + 32 * t], z);
        z = k1_warp_sum(z);
        if (lane == 0) {
          const int rr = c0 + r;
          const bool seg0 = rr < rows0;
          const bool valid = seg0 ? (rr < nv0) : true;
          const float t = seg0 ? t0 : t1;
          const float wgt = valid ? scale / (float)(seg0 ? nv0 : n1) : 0.f;
          z += smisc[0];
          const float o = act_fwd(z, p.last_act, p.slope);
          float loss, dlo;
          if (p.loss_kind == CGL_LOSS_BCE) {
            const float lo = fmaxf(logf(o), -100.f);
            const float l1o = fmaxf(logf(1.f - o), -100.f);
            loss = -(t * lo + (1.f - t) * l1o);
            dlo = (o - t) / fmaxf((1.f - o) * o, 1e-12f);
          } else {  // MSE
            const float df = o - t;
            loss = df * df;
            dlo = 2.f * df;
          }
          float d0 = dlo * wgt * act_bwd_from_out(o, p.last_act, p.slope);
          if (!valid) d0 = 0.f;
          sloss[rr] = valid ? loss : 0.f;
          sdz3[r] = d0;
        }
      } else if (lane == 0) {
        sdz3[r] = 0.f;
      }
    }
    __syncthreads();
    if (tid < K1_H2) {  // column o = tid of the chunk: last layer's weight gradient, dZ2 in place of h2, db2
      const float w3o = sw3[tid];
      for (int r = 0; r < K1_R; ++r) {
        float v = 0.f;
        if (r < nr) {
          const float hv = sH2[r * K1_LD2 + tid];
          const float dz = sdz3[r];
          if (train) gW3 = fmaf(dz, hv, gW3);
          v = (dz * w3o) * act_bwd_from_out(hv, CGL_ACT_LRELU, p.slope);
          if (train) gb2 += v;
        }
        sH2[r * K1_LD2 + tid] = v;
      }
      if (train && tid == 0)
        for (int r = 0; r < nr; ++r) gb3 += sdz3[r];
    }
    __syncthreads();
  };

  // E: this warp's share of dZ1: 16 columns (i = 16 cw + 8 nt + 2 tig, + 1; cw = warp % 8) summed over ITS half of the
  // contraction (kh = warp / 8: o in [128 kh, +128)) -- the consumers (dW1 / db1 or dXg) are linear in dZ1, the halves meet
  // there. dz1[mt][nt][e], e as the accumulator fragment (e = 0, 1: row gid; 2, 3: row gid + 8; even e: column 2 tig, odd:
  // 2 tig + 1); 0 outside the chunk. The k index of a step is permuted (slot tig <-> o = k0 + 2 tig, slot tig + 4 <->
  // k0 + 2 tig + 1: A and B agree, any order of k sums the same) so that both fragment loads are conflict-free.
  auto phase_e = [&](int nr, float (&dz1)[3][2][4]) {
    const int nm = (nr + 15) >> 4;
    const int cw = warp & 7, kh = warp >> 3;
    const int i0 = cw * 16;
#pragma unroll
    for (int mt = 0; mt < 3; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) dz1[mt][nt][e] = 0.f;
    const float* wq = sW2 + (kh * (K1_H2 / 2) + 2 * tig) * K1_LDW + i0 + gid;
    const float* zq = sH2 + gid * K1_LD2 + kh * (K1_H2 / 2) + 2 * tig;
#pragma unroll 2
    for (int k0 = 0; k0 < K1_H2 / 2; k0 += 8) {
      uint32_t bh[2][2], bl[2][2];
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const float bx[2] = {wq[k0 * K1_LDW + nt * 8], wq[(k0 + 1) * K1_LDW + nt * 8]};
        k1_split_n(bx, bh[nt], bl[nt]);
      }
#pragma unroll
      for (int mt = 0; mt < 3; ++mt) {
        if (mt < nm) {
          const float2 lo2 = *reinterpret_cast<const float2*>(zq + mt * 16 * K1_LD2 + k0);
          const float2 hi2 = *reinterpret_cast<const float2*>(zq + (mt * 16 + 8) * K1_LD2 + k0);
          const float ax[4] = {lo2.x, hi2.x, lo2.y, hi2.y};
          uint32_t ah[4], al[4];
          k1_split_n(ax, ah, al);
          k1_mma3<2>(dz1[mt], ah, al, bh, bl);
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < 3; ++mt) {
      if (mt < nm) {
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const float* hr = sH1 + (mt * 16 + gid) * K1_LD1 + i0 + nt * 8 + 2 * tig;
          const float2 h0 = *reinterpret_cast<const float2*>(hr);
          const float2 h8 = *reinterpret_cast<const float2*>(hr + 8 * K1_LD1);
          dz1[mt][nt][0] *= act_bwd_from_out(h0.x, CGL_ACT_LRELU, p.slope);
          dz1[mt][nt][1] *= act_bwd_from_out(h0.y, CGL_ACT_LRELU, p.slope);
          dz1[mt][nt][2] *= act_bwd_from_out(h8.x, CGL_ACT_LRELU, p.slope);
          dz1[mt][nt][3] *= act_bwd_from_out(h8.y, CGL_ACT_LRELU, p.slope);
        }
      }
    }
  };

  // =========================================== D step ===========================================
  if (p.do_d) {
    // dW2 tile of this warp: o in [64 mw, +64) (4 m tiles), i in [32 nw, +32) (4 n tiles)
    const int mw = warp >> 2, nw = warp & 3;
    float gW2[4][4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int e = 0; e < 4; ++e) gW2[a][b][e] = 0.f;
    float gW3 = 0.f, gb2 = 0.f, gb3 = 0.f;
    // first layer's gradients: summed per chunk over the lanes that share a column (shuffles), then kept in shared memory,
    // one private slot per (warp, tig, column): sred[((warp * 4 + tig) * 4 + j) * 3 + {0: db1, 1 + c: dW1[.][c]}]
    for (int i = tid; i < 16 * 4 * 4 * (1 + K1_MAXD); i += K1_THREADS) sred[i] = 0.f;
    const int nv0 = p.n_real ? min(p.n_real[g], B) : B;
    const float* real = p.real + (long long)g * B * d;
    const float* fake = p.fake + (long long)(p.fake_idx ? p.fake_idx[g] : g) * B * d;
    const int rows = 2 * B;
    for (int c0 = 0; c0 < rows; c0 += K1_R) {
      const int nr = min(K1_R, rows - c0);
      phase_abc(c0, nr, real, fake, B, nv0, B, 1.f, 0.f, p.d_scale, true, gW3, gb2, gb3);
      {  // E -> first layer's gradients
        float dz1[3][2][4];
        phase_e(nr, dz1);
        float xr[3][2][K1_MAXD];   // the inputs of this lane's six rows
#pragma unroll
        for (int mt = 0; mt < 3; ++mt)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int c = 0; c < K1_MAXD; ++c) xr[mt][hf][c] = sX[(mt * 16 + gid + 8 * hf) * K1_MAXD + c];
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // column 16 (warp % 8) + 8 (j / 2) + 2 tig + (j % 2)
          float vb = 0.f, vw[K1_MAXD];
#pragma unroll
          for (int c = 0; c < K1_MAXD; ++c) vw[c] = 0.f;
#pragma unroll
          for (int mt = 0; mt < 3; ++mt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              const float dz = dz1[mt][j >> 1][2 * hf + (j & 1)];
              vb += dz;
#pragma unroll
              for (int c = 0; c < K1_MAXD; ++c) vw[c] = fmaf(dz, xr[mt][hf][c], vw[c]);
            }
#pragma unroll
          for (int m = 4; m < 32; m <<= 1) {
            vb += __shfl_xor_sync(0xffffffffu, vb, m);
#pragma unroll
            for (int c = 0; c < K1_MAXD; ++c) vw[c] += __shfl_xor_sync(0xffffffffu, vw[c], m);
          }
          if (gid == 0) {
            float* dst = sred + ((warp * 4 + tig) * 4 + j) * (1 + K1_MAXD);
            dst[0] += vb;
#pragma unroll
            for (int c = 0; c < K1_MAXD; ++c) dst[1 + c] += vw[c];
          }
        }
      }
      {  // D: dW2[o][i] += sum_r dZ2[r][o] * h1[r][i]   (A = dZ2^T: m = o, k = r; B = h1: k = r, n = i)
        const int nks = ((nr + 15) >> 4) * 2;   // k steps of 8 rows that hold data (the rest of the chunk is zero)
        const float* zq = sH2 + tig * K1_LD2 + mw * 64 + gid;
        const float* hq = sH1 + tig * K1_LD1 + nw * 32 + gid;
#pragma unroll 1
        for (int ks = 0; ks < nks; ++ks) {
          const float* zr = zq + ks * 8 * K1_LD2;
          const float* hr = hq + ks * 8 * K1_LD1;
          uint32_t bh[4][2], bl[4][2];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            const float bx[2] = {hr[nt * 8], hr[4 * K1_LD1 + nt * 8]};
            k1_split_n(bx, bh[nt], bl[nt]);
          }
#pragma unroll
          for (int mt = 0; mt < 4; ++mt) {
            const float ax[4] = {zr[mt * 16], zr[mt * 16 + 8], zr[4 * K1_LD2 + mt * 16], zr[4 * K1_LD2 + mt * 16 + 8]};
            uint32_t ah[4], al[4];
            k1_split_n(ax, ah, al);
            k1_mma3<4>(gW2[mt], ah, al, bh, bl);
          }
        }
      }
      __syncthreads();   // the chunk's buffers are rewritten by the next chunk
    }

    // ---- loss (rows ascending per term, as head_kernel) ----
    if (tid == 0) {
      float l0 = 0.f, l1 = 0.f;
      for (int r = 0; r < B; ++r) l0 += sloss[r];
      for (int r = B; r < rows; ++r) l1 += sloss[r];
      float tot = 0.f;
      if (nv0 > 0) tot += l0 / (float)nv0;
      tot += l1 / (float)B;
      p.out_dloss[g] = tot * p.d_scale;
    }
    // ---- Adam on every parameter (torch.optim.Adam, IEEE sequence); the resident copies follow ----
    const AdamScalars as = *reinterpret_cast<const AdamScalars*>(smisc + 8);
    float* Mo = p.adam_m + (long long)rowid * p.ldp;
    float* Vo = p.adam_v + (long long)rowid * p.ldp;
    {  // W2: this thread's accumulator fragments (pairs of neighbouring i: 8-byte accesses, whole 32-byte sectors per row)
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        float2 m2[4][2], v2[4][2];     // the m / v of a whole m tile are requested before the first of them is used
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const long long off = p.w_off[1] + (long long)(mw * 64 + mt * 16 + gid + 8 * hf) * K1_H1 + nw * 32 + nt * 8 + 2 * tig;
            m2[nt][hf] = __ldcs(reinterpret_cast<const float2*>(Mo + off));
            v2[nt][hf] = __ldcs(reinterpret_cast<const float2*>(Vo + off));
          }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int o = mw * 64 + mt * 16 + gid + 8 * hf;
            const int i = nw * 32 + nt * 8 + 2 * tig;
            const long long off = p.w_off[1] + (long long)o * K1_H1 + i;
            float2 w2 = *reinterpret_cast<const float2*>(sW2 + o * K1_LDW + i);
            adam_update(w2.x, m2[nt][hf].x, v2[nt][hf].x, gW2[mt][nt][2 * hf + 0], as);
            adam_update(w2.y, m2[nt][hf].y, v2[nt][hf].y, gW2[mt][nt][2 * hf + 1], as);
            *reinterpret_cast<float2*>(P + off) = w2;
            __stcs(reinterpret_cast<float2*>(Mo + off), m2[nt][hf]);
            __stcs(reinterpret_cast<float2*>(Vo + off), v2[nt][hf]);
            *reinterpret_cast<float2*>(sW2 + o * K1_LDW + i) = w2;
          }
      }
    }
    if (tid < K1_H2) {  // b2 and the last layer's weights: column tid
      long long off = p.b_off[1] + tid;
      float w = sb2[tid], mm = Mo[off], vv = Vo[off];
      adam_update(w, mm, vv, gb2, as);
      P[off] = w; Mo[off] = mm; Vo[off] = vv;
      off = p.w_off[2] + tid;
      float w3 = sw3[tid];
      mm = Mo[off]; vv = Vo[off];
      adam_update(w3, mm, vv, gW3, as);
      P[off] = w3; Mo[off] = mm; Vo[off] = vv;
      sb2[tid] = w; sw3[tid] = w3;      // (every reader of the old values passed the barrier that ends the last chunk)
    }
    if (tid == 0) {
      const long long off = p.b_off[2];
      float w = smisc[0], mm = Mo[off], vv = Vo[off];
      adam_update(w, mm, vv, gb3, as);
      P[off] = w; Mo[off] = mm; Vo[off] = vv;
      smisc[0] = w;
      p.step[rowid] += 1;
    }
    if (warp < 8 && gid == 0) {  // first layer: lanes 0..3 of warps 0..7 own columns 16 warp + 8 nt + 2 tig + j
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = warp * 16 + (j >> 1) * 8 + 2 * tig + (j & 1);
        const float* lo_half = sred + ((warp * 4 + tig) * 4 + j) * (1 + K1_MAXD);          // contraction half 0: this warp
        const float* hi_half = sred + (((warp + 8) * 4 + tig) * 4 + j) * (1 + K1_MAXD);    // half 1: warp + 8
        const float gb1 = lo_half[0] + hi_half[0];
        float gW1[K1_MAXD];
#pragma unroll
        for (int c = 0; c < K1_MAXD; ++c) gW1[c] = lo_half[1 + c] + hi_half[1 + c];
        long long off = p.b_off[0] + i;
        float w = sb1[i], mm = Mo[off], vv = Vo[off];
        adam_update(w, mm, vv, gb1, as);
        P[off] = w; Mo[off] = mm; Vo[off] = vv;
        sb1[i] = w;
#pragma unroll
        for (int c = 0; c < K1_MAXD; ++c) {
          if (c < d) {
            off = p.w_off[0] + (long long)i * d + c;
            float w1 = sW1[i * d + c];
            mm = Mo[off]; vv = Vo[off];
            adam_update(w1, mm, vv, gW1[c], as);
            P[off] = w1; Mo[off] = mm; Vo[off] = vv;
            sW1[i * d + c] = w1;
          }
        }
      }
    }
    __syncthreads();
  }

  // =========================================== G loss ===========================================
  if (p.do_g) {
    const float* xg = p.xg + (long long)(p.xg_idx ? p.xg_idx[g] : g) * B * d;
    float dum0 = 0.f, dum1 = 0.f, dum2 = 0.f;
    for (int c0 = 0; c0 < B; c0 += K1_R) {
      const int nr = min(K1_R, B - c0);
      phase_abc(c0, nr, xg, xg, B, B, B, 1.f, 1.f, 1.f, false, dum0, dum1, dum2);
      if (p.out_dxg) {
        float dz1[3][2][4];
        phase_e(nr, dz1);
        // dXg[r][c] = sum_i dZ1[r][i] * W1[i][c]: this thread's four columns, the four lanes that share the row (same gid),
        // then the 16 warps (8 column groups x 2 contraction halves) through shared memory, in warp order
        float w1[4][K1_MAXD];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int c = 0; c < K1_MAXD; ++c)
            w1[j][c] = (c < d) ? sW1[((warp & 7) * 16 + (j >> 1) * 8 + 2 * tig + (j & 1)) * d + c] : 0.f;
#pragma unroll
        for (int mt = 0; mt < 3; ++mt)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf)
#pragma unroll
            for (int c = 0; c < K1_MAXD; ++c) {
              if (c < d) {
                float v = fmaf(dz1[mt][0][2 * hf + 1], w1[1][c], dz1[mt][0][2 * hf] * w1[0][c]);
                v = fmaf(dz1[mt][1][2 * hf], w1[2][c], v);
                v = fmaf(dz1[mt][1][2 * hf + 1], w1[3][c], v);
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                if (tig == 0) sred[((mt * 16 + gid + 8 * hf) * 16 + warp) * K1_MAXD + c] = v;
              }
            }
        __syncthreads();
        float* out = p.out_dxg + ((long long)g * B + c0) * d;
        for (int i = tid; i < nr * d; i += K1_THREADS) {
          const int r = i / d, c = i - r * d;
          const float* s = sred + (r * 16) * K1_MAXD + c;
          float acc = 0.f;
#pragma unroll
          for (int w = 0; w < 16; ++w) acc += s[w * K1_MAXD];
          out[i] = acc;
        }
      }
      __syncthreads();
    }
    if (tid == 0) {
      float l0 = 0.f;
      for (int r = 0; r < B; ++r) l0 += sloss[r];
      p.out_gloss[g] = l0 / (float)B;
    }
  }
}

}  // namespace cgl
