// client_fused.cuh -- K1: the whole client step of a SMALL discriminator in ONE launch, one CTA per client, with the
// client's weights resident in shared memory (north-star part 1; Worker.train, CGLGAN/2DMG/main.py:344-375 with
// D = CGLGAN/2DMG/model.py:54-71, 2 -> 128 -> 256 -> 1, 33,665 parameters = 135 KB).
//
//   D step  : X = [real | fake] (2B rows) -> h1 -> h2 -> logit -> sigmoid -> BCE (or MSE), backward through all three
//             layers, torch.optim.Adam on every parameter -- W / m / v cross HBM once (24 B per parameter), the weights
//             are read from HBM once per launch and then live in shared memory for both parts
//   G loss  : with the UPDATED weights, loss(D(Xg), valid) and dLoss/dXg (the hand-off to the server, a3)
//
// Rows are processed in chunks of K1_R = 40 (activations of a chunk live in shared memory: 60 KB next to the 135 KB of
// weights); the weight gradient of the 128 x 256 layer (98 % of the MACs) is accumulated over the chunks in REGISTERS
// (an 8 x 8 tile per thread, 512 threads), the small gradients in a few more registers per thread. Exact-fp32 FMA chains
// (contraction index ascending inside a chunk), IEEE adam_update: the arithmetic of the layered FFMA path, in another
// summation order for the gradients only (per chunk / per half of the contraction, then over those).
//
// Phases per chunk (512 threads = 16 warps, 4 per scheduler; <= 128 registers per thread):
//   A  h1 = lrelu(x W1^T + b1)                      thread: column i, every fourth row
//   B  h2 = lrelu(h1 W2^T + b2)                     thread: 10 rows x 2 columns {cg, cg + 128}; W2 rows padded to 132 floats
//   C  logit, loss term, dlogit; dW3 / db3 / db2;   warp per row, then thread per column; h2 is overwritten by dZ2
//   E  dZ1 = (dZ2 W2) * lrelu'(h1)                  thread: 10 rows x 2 columns {cg, cg + 64} over HALF of the contraction (the
//                                                   consumers -- dW1 / db1 or dXg -- are linear in dZ1: the halves meet there)
//   D  dW2 += dZ2^T h1                              thread: 8 x 8 tile, one row of the chunk per iteration
// A first version with 256 threads and a 16 x 8 tile (246 registers) ran at 35 % of the FFMA peak: with two warps per
// scheduler half of the issue slots stayed empty behind shared-memory latency (profiles/ncu_k1_r2.md).
#pragma once
#include "linear.cuh"

namespace cgl {

constexpr int K1_H1 = 128;
constexpr int K1_H2 = 256;
constexpr int K1_R = 40;
constexpr int K1_THREADS = 512;
constexpr int K1_LDW = 132;     // floats per row of the resident W2 copy (conflict-free float4 reads down a column of rows)
constexpr int K1_MAXD = 2;      // widest input
constexpr int K1_MAXROWS = 512; // 2B
#ifndef K1_DEFAULT_ON
#define K1_DEFAULT_ON 0      // measured on B200: 1.99 ms against 1.47 ms of the layered tcgen05 kernels per 1024-client step
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

constexpr int K1_SRED = 8 * K1_H1 * (1 + K1_MAXD);   // 8 partial sums (2 contraction halves x 4 row groups) of db1 | dW1
constexpr int K1_SMEM_FLOATS = K1_H2 * K1_LDW + K1_R * K1_H1 + K1_R * K1_H2 + K1_R * K1_MAXD + K1_H1 * K1_MAXD + K1_H1 +
                               K1_H2 + K1_H2 + K1_R + K1_MAXROWS + K1_SRED + 16;
constexpr size_t K1_SMEM_BYTES = (size_t)K1_SMEM_FLOATS * sizeof(float);

__device__ __forceinline__ float k1_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(K1_THREADS, 1) client_step_fused_kernel(const K1Params p) {
  extern __shared__ __align__(16) float k1s[];
  float* sW2 = k1s;                               // [256][132]
  float* sH1 = sW2 + K1_H2 * K1_LDW;              // [R][128]
  float* sH2 = sH1 + K1_R * K1_H1;                // [R][256]   h2, then dZ2
  float* sX = sH2 + K1_R * K1_H2;                 // [R][2]
  float* sW1 = sX + K1_R * K1_MAXD;               // [128][d]
  float* sb1 = sW1 + K1_H1 * K1_MAXD;             // [128]
  float* sb2 = sb1 + K1_H1;                       // [256]
  float* sw3 = sb2 + K1_H2;                       // [256]
  float* sdz3 = sw3 + K1_H2;                      // [R]
  float* sloss = sdz3 + K1_R;                     // [2B]
  float* sred = sloss + K1_MAXROWS;               // [8][128][1 + d]  (also [R][4][d] for dXg)
  float* smisc = sred + K1_SRED;                  // [0] = b3, [8..16) = AdamScalars of this step

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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

  // thread coordinates of the phases
  const int rgB = tid >> 7, cgB = tid & 127;                // B: row group (10 rows), columns cgB, cgB + 128
  const int ks = tid >> 8, rgE = (tid >> 6) & 3, cgE = tid & 63;   // E: contraction half, row group, columns cgE, cgE + 64
  const int og = tid >> 4, ig = tid & 15;                   // D: 8 o x (4 + 4) i

  // ---- one chunk of rows through phases A, B, C (shared by both parts) ----
  // src rows: rr < rows0 -> p0 + rr * d, else p1 + (rr - rows0) * d
  auto phase_abc = [&](int c0, int nr, const float* p0, const float* p1, int rows0, int nv0, int n1, float t0, float t1,
                       float scale, bool train, float& gW3, float& gb2, float& gb3) {
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
        sH1[r * K1_H1 + i] = h;
      }
    }
    __syncthreads();
    if (rgB * 10 < nr) {  // B
      float acc[10][2];
#pragma unroll
      for (int r = 0; r < 10; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; }
      const float* h = sH1 + rgB * 10 * K1_H1;
      const float* w = sW2 + cgB * K1_LDW;
#pragma unroll 2
      for (int k = 0; k < K1_H1; k += 4) {
        const float4 w0 = *reinterpret_cast<const float4*>(w + k);
        const float4 w1 = *reinterpret_cast<const float4*>(w + 128 * K1_LDW + k);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(h + r * K1_H1 + k);
          acc[r][0] = fmaf(a.x, w0.x, acc[r][0]);
          acc[r][0] = fmaf(a.y, w0.y, acc[r][0]);
          acc[r][0] = fmaf(a.z, w0.z, acc[r][0]);
          acc[r][0] = fmaf(a.w, w0.w, acc[r][0]);
          acc[r][1] = fmaf(a.x, w1.x, acc[r][1]);
          acc[r][1] = fmaf(a.y, w1.y, acc[r][1]);
          acc[r][1] = fmaf(a.z, w1.z, acc[r][1]);
          acc[r][1] = fmaf(a.w, w1.w, acc[r][1]);
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int o = cgB + 128 * j;
        const float bb = sb2[o];
#pragma unroll
        for (int r = 0; r < 10; ++r)
          sH2[(rgB * 10 + r) * K1_H2 + o] = act_fwd(acc[r][j] + bb, CGL_ACT_LRELU, p.slope);
      }
    }
    __syncthreads();
    // C: logits, loss terms, d loss / d logit (the arithmetic of head_kernel, dstep.cu)
    for (int r = warp; r < K1_R; r += K1_THREADS / 32) {
      if (r < nr) {
        float z = 0.f;
#pragma unroll
        for (int t = 0; t < K1_H2 / 32; ++t) z = fmaf(sH2[r * K1_H2 + lane + 32 * t], sw3[lane + 32 * t], z);
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
          const float hv = sH2[r * K1_H2 + tid];
          const float dz = sdz3[r];
          if (train) gW3 = fmaf(dz, hv, gW3);
          v = (dz * w3o) * act_bwd_from_out(hv, CGL_ACT_LRELU, p.slope);
          if (train) gb2 += v;
        }
        sH2[r * K1_H2 + tid] = v;
      }
      if (train && tid == 0)
        for (int r = 0; r < nr; ++r) gb3 += sdz3[r];
    }
    __syncthreads();
  };

  // E: this thread's share of dZ1 -- 10 rows x 2 columns, summed over ITS half of the contraction (0 outside the chunk)
  auto phase_e = [&](int nr, float (&dz1)[10][2]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) { dz1[r][0] = 0.f; dz1[r][1] = 0.f; }
    if (rgE * 10 >= nr) return;
    const float* dz = sH2 + rgE * 10 * K1_H2 + ks * (K1_H2 / 2);
    const float* w = sW2 + (ks * (K1_H2 / 2)) * K1_LDW + cgE;
#pragma unroll 2
    for (int o = 0; o < K1_H2 / 2; o += 4) {
      float wv[4][2];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        wv[e][0] = w[(o + e) * K1_LDW];
        wv[e][1] = w[(o + e) * K1_LDW + 64];
      }
#pragma unroll
      for (int r = 0; r < 10; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(dz + r * K1_H2 + o);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          dz1[r][j] = fmaf(a.x, wv[0][j], dz1[r][j]);
          dz1[r][j] = fmaf(a.y, wv[1][j], dz1[r][j]);
          dz1[r][j] = fmaf(a.z, wv[2][j], dz1[r][j]);
          dz1[r][j] = fmaf(a.w, wv[3][j], dz1[r][j]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 10; ++r)
#pragma unroll
      for (int j = 0; j < 2; ++j)
        dz1[r][j] *= act_bwd_from_out(sH1[(rgE * 10 + r) * K1_H1 + cgE + 64 * j], CGL_ACT_LRELU, p.slope);
  };

  // =========================================== D step ===========================================
  if (p.do_d) {
    float gW2[8][8];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 8; ++b) gW2[a][b] = 0.f;
    float gW3 = 0.f, gb2 = 0.f, gb3 = 0.f;
    float gb1[2] = {0.f, 0.f};
    float gW1[2][K1_MAXD];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < K1_MAXD; ++c) gW1[j][c] = 0.f;

    const int nv0 = p.n_real ? min(p.n_real[g], B) : B;
    const float* real = p.real + (long long)g * B * d;
    const float* fake = p.fake + (long long)(p.fake_idx ? p.fake_idx[g] : g) * B * d;
    const int rows = 2 * B;
    for (int c0 = 0; c0 < rows; c0 += K1_R) {
      const int nr = min(K1_R, rows - c0);
      phase_abc(c0, nr, real, fake, B, nv0, B, 1.f, 0.f, p.d_scale, true, gW3, gb2, gb3);
      {  // E -> first layer's gradients (this thread's 10 rows of 2 columns, its half of the contraction)
        float dz1[10][2];
        phase_e(nr, dz1);
#pragma unroll
        for (int r = 0; r < 10; ++r)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            gb1[j] += dz1[r][j];
#pragma unroll
            for (int c = 0; c < K1_MAXD; ++c)
              if (c < d) gW1[j][c] = fmaf(dz1[r][j], sX[(rgE * 10 + r) * K1_MAXD + c], gW1[j][c]);
          }
      }
      {  // D: dW2[o][i] += dZ2[r][o] * h1[r][i]
        const float* dz = sH2 + og * 8;
        const float* h = sH1 + 4 * ig;
#pragma unroll 2
        for (int r = 0; r < nr; ++r) {
          const float4 ha = *reinterpret_cast<const float4*>(h + r * K1_H1);
          const float4 hb = *reinterpret_cast<const float4*>(h + r * K1_H1 + 64);
          const float hv[8] = {ha.x, ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w};
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float4 z4 = *reinterpret_cast<const float4*>(dz + r * K1_H2 + 4 * q);
            const float zv[4] = {z4.x, z4.y, z4.z, z4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
#pragma unroll
              for (int b = 0; b < 8; ++b) gW2[4 * q + e][b] = fmaf(zv[e], hv[b], gW2[4 * q + e][b]);
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
    // ---- first layer's gradients: the eight partial sums (contraction half x row group) meet in shared memory ----
    {
      const int st = 1 + d;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float* dst = sred + ((ks * 4 + rgE) * K1_H1 + cgE + 64 * j) * st;
        dst[0] = gb1[j];
#pragma unroll
        for (int c = 0; c < K1_MAXD; ++c)
          if (c < d) dst[1 + c] = gW1[j][c];
      }
    }
    __syncthreads();

    // ---- Adam on every parameter (torch.optim.Adam, IEEE sequence); the resident copies follow ----
    const AdamScalars as = *reinterpret_cast<const AdamScalars*>(smisc + 8);
    float* Mo = p.adam_m + (long long)rowid * p.ldp;
    float* Vo = p.adam_v + (long long)rowid * p.ldp;
    {  // W2: this thread's 8 x 8 tile (fully unrolled: the accumulators are registers)
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int o = og * 8 + a;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int i0 = hf * 64 + 4 * ig;
          const long long off = p.w_off[1] + (long long)o * K1_H1 + i0;
          float4 w4 = *reinterpret_cast<const float4*>(sW2 + o * K1_LDW + i0);
          float4 m4 = *reinterpret_cast<const float4*>(Mo + off);
          float4 v4 = *reinterpret_cast<const float4*>(Vo + off);
          adam_update(w4.x, m4.x, v4.x, gW2[a][4 * hf + 0], as);
          adam_update(w4.y, m4.y, v4.y, gW2[a][4 * hf + 1], as);
          adam_update(w4.z, m4.z, v4.z, gW2[a][4 * hf + 2], as);
          adam_update(w4.w, m4.w, v4.w, gW2[a][4 * hf + 3], as);
          *reinterpret_cast<float4*>(P + off) = w4;
          *reinterpret_cast<float4*>(Mo + off) = m4;
          *reinterpret_cast<float4*>(Vo + off) = v4;
          *reinterpret_cast<float4*>(sW2 + o * K1_LDW + i0) = w4;
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
    if (tid >= K1_THREADS - K1_H1) {  // first layer (the last four warps): partial sums in a fixed order
      const int i = tid - (K1_THREADS - K1_H1);
      const int st = 1 + d;
      float gb = 0.f, gw[K1_MAXD];
#pragma unroll
      for (int c = 0; c < K1_MAXD; ++c) gw[c] = 0.f;
      for (int q = 0; q < 8; ++q) {
        const float* src = sred + (q * K1_H1 + i) * st;
        gb += src[0];
#pragma unroll
        for (int c = 0; c < K1_MAXD; ++c)
          if (c < d) gw[c] += src[1 + c];
      }
      long long off = p.b_off[0] + i;
      float w = sb1[i], mm = Mo[off], vv = Vo[off];
      adam_update(w, mm, vv, gb, as);
      P[off] = w; Mo[off] = mm; Vo[off] = vv;
      sb1[i] = w;
#pragma unroll
      for (int c = 0; c < K1_MAXD; ++c) {
        if (c < d) {
          off = p.w_off[0] + (long long)i * d + c;
          float w1 = sW1[i * d + c];
          mm = Mo[off]; vv = Vo[off];
          adam_update(w1, mm, vv, gw[c], as);
          P[off] = w1; Mo[off] = mm; Vo[off] = vv;
          sW1[i * d + c] = w1;
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
        float dz1[10][2];
        phase_e(nr, dz1);
        // dXg[r][c] = sum_i dZ1[r][i] * W1[i][c]: this thread's two columns and half contraction, then the 128 threads
        // (4 warps: contraction half x column half) that share the row
        float w1[2][K1_MAXD];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
          for (int c = 0; c < K1_MAXD; ++c) w1[j][c] = (c < d) ? sW1[(cgE + 64 * j) * d + c] : 0.f;
        const int part = ks * 2 + (warp & 1);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
#pragma unroll
          for (int c = 0; c < K1_MAXD; ++c) {
            if (c < d) {
              float v = fmaf(dz1[r][1], w1[1][c], dz1[r][0] * w1[0][c]);
              v = k1_warp_sum(v);
              if (lane == 0) sred[((rgE * 10 + r) * 4 + part) * K1_MAXD + c] = v;
            }
          }
        }
        __syncthreads();
        float* out = p.out_dxg + ((long long)g * B + c0) * d;
        for (int i = tid; i < nr * d; i += K1_THREADS) {
          const int r = i / d, c = i - r * d;
          const float* s = sred + (r * 4) * K1_MAXD + c;
          out[i] = (s[0] + s[K1_MAXD]) + (s[2 * K1_MAXD] + s[3 * K1_MAXD]);
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
