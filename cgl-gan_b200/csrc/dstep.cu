// dstep.cu -- the per-client discriminator step and generator-loss evaluation (K1/K2).
// Reference: Worker.train, CGLGAN/2DMG/main.py:344-375; capgan.py:316-349; MDGAN/MNIST/mdgan.py:266-297.
#include "linear.cuh"
#include "client_fused.cuh"
#include <stdlib.h>

namespace cgl {

// CGL_GEMM_MODE=0|1|2 in the environment presets cgl_set_gemm_mode (profiling and bisecting only)
static int initial_gemm_mode() {
  const char* e = getenv("CGL_GEMM_MODE");
  if (e && e[0] >= '0' && e[0] <= '2' && e[1] == 0) return e[0] - '0';
  return GEMM_AUTO;
}
static int g_gemm_mode = initial_gemm_mode();
int gemm_mode() { return g_gemm_mode; }

// ---------------------------------------------------------------------------------------------
// Head kernel: last Linear (H -> nout, nout <= 2) + output activation + loss + its backward.
// One CTA per group. Produces the loss, dZ of the last hidden layer, and (train) the Adam update
// of the last layer.  BCE/CE/MSE follow torch.nn semantics (mean reduction per term; BCE log
// clamped at -100; CE = log_softmax + nll).
// ---------------------------------------------------------------------------------------------
struct HeadParams {
  int H, nout, rows, rows0;
  const float* hin;  // [G][rows][H] last hidden activations
  long long hin_gstride;
  float* dz;  // [G][rows][H] out: grad wrt pre-activation of the last hidden layer
  long long dz_gstride;
  int hidden_act;
  float slope;
  float* params; float* adam_m; float* adam_v;  // packed rows (adam_* only when train)
  long long ldp;
  const int* ids;
  long long w_off, b_off;
  const int* step;
  float lr, b1, b2, eps;
  const AdamScalars* scal;  // [G] (train)
  const int* n_valid0;  // [G] valid rows among the first rows0 (NULL: all)
  int loss_kind, last_act;
  float target0, target1;  // targets of rows [0,rows0) and [rows0,rows)
  float scale;
  float* out_loss;  // [G]
  int train;
};

constexpr int HEAD_THREADS = 256;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(HEAD_THREADS) head_kernel(const HeadParams p) {
  extern __shared__ float smem[];
  float* sW = smem;                        // [nout][H]
  float* sdz = sW + p.nout * p.H;          // [rows][nout]
  float* sloss = sdz + p.rows * p.nout;    // [rows]
  float* sb = sloss + p.rows;              // [nout]

  const int g = blockIdx.x;
  const int rowid = p.ids ? p.ids[g] : g;
  float* W = p.params + (long long)rowid * p.ldp + p.w_off;
  float* Bv = p.params + (long long)rowid * p.ldp + p.b_off;
  const float* hin = p.hin + (long long)g * p.hin_gstride;
  float* dz = p.dz + (long long)g * p.dz_gstride;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarp = HEAD_THREADS / 32;
  const int H = p.H, nout = p.nout;

  for (int i = tid; i < nout * H; i += HEAD_THREADS) sW[i] = W[i];
  if (tid < nout) sb[tid] = Bv[tid];
  __syncthreads();

  const int nv0 = p.n_valid0 ? min(p.n_valid0[g], p.rows0) : p.rows0;
  const int n1 = p.rows - p.rows0;

  // ---- logits, loss terms, d loss / d logits ----
  for (int r = warp; r < p.rows; r += nwarp) {
    const float* hr = hin + (long long)r * H;
    float z0 = 0.f, z1 = 0.f;
    for (int h = lane; h < H; h += 32) {
      float a = hr[h];
      z0 = fmaf(a, sW[h], z0);
      if (nout == 2) z1 = fmaf(a, sW[H + h], z1);
    }
    z0 = warp_sum(z0);
    if (nout == 2) z1 = warp_sum(z1);
    if (lane == 0) {
      const bool seg0 = r < p.rows0;
      const bool valid = seg0 ? (r < nv0) : true;
      const float t = seg0 ? p.target0 : p.target1;
      const float wgt = valid ? p.scale / (float)(seg0 ? nv0 : n1) : 0.f;
      z0 += sb[0];
      if (nout == 2) z1 += sb[1];
      float loss = 0.f, d0 = 0.f, d1 = 0.f;
      if (p.loss_kind == CGL_LOSS_CE) {
        // log_softmax + nll_loss, integer class target = (int)t
        float mx = fmaxf(z0, z1);
        float lse = mx + logf(expf(z0 - mx) + expf(z1 - mx));
        int cls = (int)t;
        loss = lse - (cls == 0 ? z0 : z1);
        float s0 = expf(z0 - lse), s1 = expf(z1 - lse);
        d0 = (s0 - (cls == 0 ? 1.f : 0.f)) * wgt;
        d1 = (s1 - (cls == 1 ? 1.f : 0.f)) * wgt;
      } else {
        float o = act_fwd(z0, p.last_act, p.slope);
        float dlo;  // d loss / d o
        if (p.loss_kind == CGL_LOSS_BCE) {
          float lo = fmaxf(logf(o), -100.f);
          float l1o = fmaxf(logf(1.f - o), -100.f);
          loss = -(t * lo + (1.f - t) * l1o);
          dlo = (o - t) / fmaxf((1.f - o) * o, 1e-12f);
        } else {  // MSE
          float d = o - t;
          loss = d * d;
          dlo = 2.f * d;
        }
        d0 = dlo * wgt * act_bwd_from_out(o, p.last_act, p.slope);
      }
      if (!valid) { d0 = 0.f; d1 = 0.f; }  // padded rows of a ragged real batch carry no gradient
      sloss[r] = valid ? loss : 0.f;
      sdz[r * nout] = d0;
      if (nout == 2) sdz[r * nout + 1] = d1;
    }
  }
  __syncthreads();

  if (tid == 0) {
    float l0 = 0.f, l1 = 0.f;
    for (int r = 0; r < p.rows0; ++r) l0 += sloss[r];
    for (int r = p.rows0; r < p.rows; ++r) l1 += sloss[r];
    float tot = 0.f;
    if (nv0 > 0) tot += l0 / (float)nv0;
    if (n1 > 0) tot += l1 / (float)n1;
    p.out_loss[g] = tot * p.scale;
  }

  // ---- dZ of the last hidden layer: (dlogits . W) * act'(h) ----
  for (int idx = tid; idx < p.rows * H; idx += HEAD_THREADS) {
    int r = idx / H, h = idx - r * H;
    float a = hin[idx];
    float v = sdz[r * nout] * sW[h];
    if (nout == 2) v = fmaf(sdz[r * nout + 1], sW[H + h], v);
    dz[idx] = v * act_bwd_from_out(a, p.hidden_act, p.slope);
  }

  // ---- Adam on the last layer: dW[j][h] = sum_r dlogit[r][j] * h[r][h]; db[j] = sum_r dlogit[r][j]
  if (p.train) {
    const AdamScalars s = p.scal ? p.scal[g] : make_adam_scalars(p.step[rowid], p.lr, p.b1, p.b2, p.eps);
    float* Mo = p.adam_m + (long long)rowid * p.ldp;
    float* Vo = p.adam_v + (long long)rowid * p.ldp;
    for (int h = tid; h < H; h += HEAD_THREADS) {
      float g0 = 0.f, g1 = 0.f;
      for (int r = 0; r < p.rows; ++r) {
        float a = hin[(long long)r * H + h];
        g0 = fmaf(sdz[r * nout], a, g0);
        if (nout == 2) g1 = fmaf(sdz[r * nout + 1], a, g1);
      }
      {
        long long o = p.w_off + h;
        float w = sW[h], mm = Mo[o], vv = Vo[o];
        adam_update(w, mm, vv, g0, s);
        W[h] = w; Mo[o] = mm; Vo[o] = vv;
      }
      if (nout == 2) {
        long long o = p.w_off + H + h;
        float w = sW[H + h], mm = Mo[o], vv = Vo[o];
        adam_update(w, mm, vv, g1, s);
        W[H + h] = w; Mo[o] = mm; Vo[o] = vv;
      }
    }
    if (tid < nout) {
      float gb = 0.f;
      for (int r = 0; r < p.rows; ++r) gb += sdz[r * nout + tid];
      long long o = p.b_off + tid;
      float w = sb[tid], mm = Mo[o], vv = Vo[o];
      adam_update(w, mm, vv, gb, s);
      Bv[tid] = w; Mo[o] = mm; Vo[o] = vv;
    }
  }
}

// The same head as ONE pass over the last hidden layer: the activations travel global -> shared in chunks of
// HEAD_CH rows (16-byte cp.async, double-buffered: the next chunk is in flight while this one is consumed), and the
// three consumers of a row -- its logits, dZ of the hidden layer, and the last layer's weight gradient -- read it
// from there. head_kernel reads the 200 x 256 block three times with dependent 4-byte loads (1.4 TB/s); the
// arithmetic and its order are unchanged (rows ascending, the same warp reduction), so the results are bit-identical.
// Needs H % 4 == 0, H <= 1024 (weight-gradient accumulators stay in registers) and float4-addressable rows.
constexpr int HEAD_CH = 50;
constexpr int HEAD_MAXJ = 4;   // H / HEAD_THREADS

__global__ void __launch_bounds__(HEAD_THREADS) head_stream_kernel(const HeadParams p) {
  extern __shared__ __align__(16) float smem[];
  const int H = p.H, nout = p.nout;
  float* tile0 = smem;                       // [HEAD_CH][H]
  float* tile1 = tile0 + HEAD_CH * H;        // [HEAD_CH][H]
  float* sW = tile1 + HEAD_CH * H;           // [nout][H]
  float* sdz = sW + nout * H;                // [rows][nout]
  float* sloss = sdz + p.rows * nout;        // [rows]
  float* sb = sloss + p.rows;                // [nout]

  const int g = blockIdx.x;
  const int rowid = p.ids ? p.ids[g] : g;
  float* W = p.params + (long long)rowid * p.ldp + p.w_off;
  float* Bv = p.params + (long long)rowid * p.ldp + p.b_off;
  const float* hin = p.hin + (long long)g * p.hin_gstride;
  float* dz = p.dz + (long long)g * p.dz_gstride;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarp = HEAD_THREADS / 32;
  const int nchunks = (p.rows + HEAD_CH - 1) / HEAD_CH;

  auto issue_chunk = [&](int c) {
    float* T = (c & 1) ? tile1 : tile0;
    const int r0 = c * HEAD_CH;
    const int nr = min(HEAD_CH, p.rows - r0);
    const float* src = hin + (long long)r0 * H;
    for (int i = tid; i < nr * (H >> 2); i += HEAD_THREADS)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(T + 4 * i)),
                   "l"(src + 4 * i) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  issue_chunk(0);
  for (int i = tid; i < nout * H; i += HEAD_THREADS) sW[i] = W[i];
  if (tid < nout) sb[tid] = Bv[tid];

  const int nv0 = p.n_valid0 ? min(p.n_valid0[g], p.rows0) : p.rows0;
  const int n1 = p.rows - p.rows0;
  float g0[HEAD_MAXJ], g1[HEAD_MAXJ];
#pragma unroll
  for (int j = 0; j < HEAD_MAXJ; ++j) { g0[j] = 0.f; g1[j] = 0.f; }

  for (int c = 0; c < nchunks; ++c) {
    if (c + 1 < nchunks) {
      issue_chunk(c + 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();                         // chunk c (and, the first time, sW / sb) visible to every thread
    const float* T = (c & 1) ? tile1 : tile0;
    const int r0 = c * HEAD_CH;
    const int nr = min(HEAD_CH, p.rows - r0);

    // ---- logits, loss terms, d loss / d logits of the chunk's rows ----
    for (int r = r0 + warp; r < r0 + nr; r += nwarp) {
      const float* hr = T + (r - r0) * H;
      float z0 = 0.f, z1 = 0.f;
      for (int h = lane; h < H; h += 32) {
        float a = hr[h];
        z0 = fmaf(a, sW[h], z0);
        if (nout == 2) z1 = fmaf(a, sW[H + h], z1);
      }
      z0 = warp_sum(z0);
      if (nout == 2) z1 = warp_sum(z1);
      if (lane == 0) {
        const bool seg0 = r < p.rows0;
        const bool valid = seg0 ? (r < nv0) : true;
        const float t = seg0 ? p.target0 : p.target1;
        const float wgt = valid ? p.scale / (float)(seg0 ? nv0 : n1) : 0.f;
        z0 += sb[0];
        if (nout == 2) z1 += sb[1];
        float loss = 0.f, d0 = 0.f, d1 = 0.f;
        if (p.loss_kind == CGL_LOSS_CE) {
          float mx = fmaxf(z0, z1);
          float lse = mx + logf(expf(z0 - mx) + expf(z1 - mx));
          int cls = (int)t;
          loss = lse - (cls == 0 ? z0 : z1);
          float s0 = expf(z0 - lse), s1 = expf(z1 - lse);
          d0 = (s0 - (cls == 0 ? 1.f : 0.f)) * wgt;
          d1 = (s1 - (cls == 1 ? 1.f : 0.f)) * wgt;
        } else {
          float o = act_fwd(z0, p.last_act, p.slope);
          float dlo;
          if (p.loss_kind == CGL_LOSS_BCE) {
            float lo = fmaxf(logf(o), -100.f);
            float l1o = fmaxf(logf(1.f - o), -100.f);
            loss = -(t * lo + (1.f - t) * l1o);
            dlo = (o - t) / fmaxf((1.f - o) * o, 1e-12f);
          } else {
            float d = o - t;
            loss = d * d;
            dlo = 2.f * d;
          }
          d0 = dlo * wgt * act_bwd_from_out(o, p.last_act, p.slope);
        }
        if (!valid) { d0 = 0.f; d1 = 0.f; }
        sloss[r] = valid ? loss : 0.f;
        sdz[r * nout] = d0;
        if (nout == 2) sdz[r * nout + 1] = d1;
      }
    }
    __syncthreads();

    // ---- dZ of the last hidden layer for the chunk: (dlogits . W) * act'(h), four consecutive h per thread ----
    for (int i = tid; i < nr * (H >> 2); i += HEAD_THREADS) {
      const int rl = i / (H >> 2), h = 4 * (i - rl * (H >> 2));
      const int r = r0 + rl;
      const float4 a4 = *reinterpret_cast<const float4*>(T + rl * H + h);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      float ov[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float v = sdz[r * nout] * sW[h + e];
        if (nout == 2) v = fmaf(sdz[r * nout + 1], sW[H + h + e], v);
        ov[e] = v * act_bwd_from_out(av[e], p.hidden_act, p.slope);
      }
      *reinterpret_cast<float4*>(dz + (long long)r * H + h) = make_float4(ov[0], ov[1], ov[2], ov[3]);
    }

    // ---- weight gradient of the last layer, rows ascending (continued over the chunks) ----
    if (p.train) {
#pragma unroll
      for (int j = 0; j < HEAD_MAXJ; ++j) {
        const int h = tid + j * HEAD_THREADS;
        if (h < H) {
          for (int rl = 0; rl < nr; ++rl) {
            const float a = T[rl * H + h];
            g0[j] = fmaf(sdz[(r0 + rl) * nout], a, g0[j]);
            if (nout == 2) g1[j] = fmaf(sdz[(r0 + rl) * nout + 1], a, g1[j]);
          }
        }
      }
    }
    __syncthreads();                         // the tile is overwritten by the chunk requested in the next iteration
  }

  if (tid == 0) {
    float l0 = 0.f, l1 = 0.f;
    for (int r = 0; r < p.rows0; ++r) l0 += sloss[r];
    for (int r = p.rows0; r < p.rows; ++r) l1 += sloss[r];
    float tot = 0.f;
    if (nv0 > 0) tot += l0 / (float)nv0;
    if (n1 > 0) tot += l1 / (float)n1;
    p.out_loss[g] = tot * p.scale;
  }

  if (p.train) {
    const AdamScalars s = p.scal ? p.scal[g] : make_adam_scalars(p.step[rowid], p.lr, p.b1, p.b2, p.eps);
    float* Mo = p.adam_m + (long long)rowid * p.ldp;
    float* Vo = p.adam_v + (long long)rowid * p.ldp;
#pragma unroll
    for (int j = 0; j < HEAD_MAXJ; ++j) {
      const int h = tid + j * HEAD_THREADS;
      if (h < H) {
        {
          long long o = p.w_off + h;
          float w = sW[h], mm = Mo[o], vv = Vo[o];
          adam_update(w, mm, vv, g0[j], s);
          W[h] = w; Mo[o] = mm; Vo[o] = vv;
        }
        if (nout == 2) {
          long long o = p.w_off + H + h;
          float w = sW[H + h], mm = Mo[o], vv = Vo[o];
          adam_update(w, mm, vv, g1[j], s);
          W[H + h] = w; Mo[o] = mm; Vo[o] = vv;
        }
      }
    }
    if (tid < nout) {
      float gb = 0.f;
      for (int r = 0; r < p.rows; ++r) gb += sdz[r * nout + tid];
      long long o = p.b_off + tid;
      float w = sb[tid], mm = Mo[o], vv = Vo[o];
      adam_update(w, mm, vv, gb, s);
      Bv[tid] = w; Mo[o] = mm; Vo[o] = vv;
    }
  }
}

// ---------------------------------------------------------------------------------------------
static int validate_d_arch(const cgl_mlp_desc* a, int loss_kind) {
  CGL_REQUIRE(a != nullptr, "arch is NULL");
  CGL_REQUIRE(a->n_layers >= 2 && a->n_layers <= CGL_MAX_LAYERS, "discriminator needs 2..%d Linear layers, got %d",
              CGL_MAX_LAYERS, a->n_layers);
  for (int i = 0; i < a->n_layers; ++i) {
    CGL_REQUIRE(a->dims[i] > 0 && a->dims[i + 1] > 0, "bad layer width at layer %d", i);
    CGL_REQUIRE(a->bn[i] == 0, "BatchNorm is not supported inside a discriminator (layer %d)", i);
  }
  const int nout = a->dims[a->n_layers];
  const int last = a->act[a->n_layers - 1];
  switch (loss_kind) {
    case CGL_LOSS_BCE:
      CGL_REQUIRE(nout == 1 && last == CGL_ACT_SIGMOID, "BCE needs a 1-logit sigmoid discriminator (reference pairing, SURVEY 3.5.5)");
      break;
    case CGL_LOSS_CE:
      CGL_REQUIRE(nout == 2 && last == CGL_ACT_NONE, "CrossEntropy needs a 2-logit discriminator without output activation");
      break;
    case CGL_LOSS_MSE:
      CGL_REQUIRE(nout == 1 && (last == CGL_ACT_NONE || last == CGL_ACT_SIGMOID), "MSE needs a 1-output discriminator");
      break;
    default:
      set_error("unknown loss kind %d", loss_kind);
      return CGL_EINVAL;
  }
  CGL_REQUIRE(a->dims[a->n_layers - 1] <= 4096, "last hidden width too large for the head kernel");
  return CGL_OK;
}

static size_t acts_floats(const cgl_mlp_desc* a, int G, int rows) {
  size_t n = 0;
  for (int i = 1; i < a->n_layers; ++i) n += (size_t)G * rows * a->dims[i];
  return n;
}
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct Workspace {
  float* H[CGL_MAX_LAYERS];   // H[i], i=1..L-1 : [G][rows][dims[i]]
  float* dZ[CGL_MAX_LAYERS];  // same shapes
  AdamScalars* scal;          // [G]
};
static Workspace carve(const cgl_mlp_desc* a, int G, int rows, void* ws) {
  Workspace w;
  char* p = (char*)ws;
  for (int i = 1; i < a->n_layers; ++i) {
    size_t bytes = align_up((size_t)G * rows * a->dims[i] * sizeof(float), 256);
    w.H[i] = (float*)p; p += bytes;
    w.dZ[i] = (float*)p; p += bytes;
  }
  w.scal = (AdamScalars*)p;
  return w;
}
static size_t ws_bytes(const cgl_mlp_desc* a, int G, int rows) {
  size_t b = 0;
  for (int i = 1; i < a->n_layers; ++i) b += 2 * align_up((size_t)G * rows * a->dims[i] * sizeof(float), 256);
  return b + align_up((size_t)G * sizeof(AdamScalars), 256) + 256;
}

static int forward_hidden(const cgl_mlp_desc* a, const cgl_mlp_layout& lay, int G, int rows, const RowMap& X,
                          const float* params, long long ldp, const int* ids, Workspace& w, cudaStream_t st) {
  for (int l = 0; l + 1 < a->n_layers; ++l) {
    RowMap A = (l == 0) ? X : single_rows(w.H[l], (long long)rows * a->dims[l], nullptr, a->dims[l]);
    CGL_CHECK_CUDA(run_linear_fwd(G, rows, a->dims[l], a->dims[l + 1], A, params, ldp, ids, lay.w_off[l], lay.b_off[l],
                                  a->act[l], a->lrelu_slope, w.H[l + 1], (long long)rows * a->dims[l + 1], st));
  }
  return CGL_OK;
}

static int launch_head(const cgl_mlp_desc* a, const cgl_mlp_layout& lay, int G, int rows, int rows0, Workspace& w,
                       float* params, float* am, float* av, long long ldp, const int* ids, const int* step,
                       const cgl_train_cfg* cfg, int loss_kind, float scale, const int* n_valid0, float t0, float t1,
                       float* out_loss, int train, cudaStream_t st) {
  const int L = a->n_layers;
  HeadParams h = {};
  h.H = a->dims[L - 1]; h.nout = a->dims[L]; h.rows = rows; h.rows0 = rows0;
  h.hin = w.H[L - 1]; h.hin_gstride = (long long)rows * h.H;
  h.dz = w.dZ[L - 1]; h.dz_gstride = (long long)rows * h.H;
  h.hidden_act = a->act[L - 2]; h.slope = a->lrelu_slope;
  h.params = params; h.adam_m = am; h.adam_v = av; h.ldp = ldp; h.ids = ids;
  h.w_off = lay.w_off[L - 1]; h.b_off = lay.b_off[L - 1];
  h.step = step;
  h.scal = train ? w.scal : nullptr;
  if (cfg) { h.lr = cfg->lr; h.b1 = cfg->beta1; h.b2 = cfg->beta2; h.eps = cfg->eps; }
  h.n_valid0 = n_valid0;
  h.loss_kind = loss_kind; h.last_act = a->act[L - 1];
  h.target0 = t0; h.target1 = t1; h.scale = scale;
  h.out_loss = out_loss; h.train = train;
  size_t smem = (size_t)(h.nout * h.H + rows * h.nout + rows + h.nout) * sizeof(float);
  const size_t smem_stream = smem + (size_t)2 * HEAD_CH * h.H * sizeof(float);
  const bool stream_ok = h.H % 4 == 0 && h.H <= HEAD_MAXJ * HEAD_THREADS && aligned16(h.hin) && aligned16(h.dz) &&
                         h.hin_gstride % 4 == 0 && h.dz_gstride % 4 == 0 && smem_stream <= 200 * 1024;
  ProfScope prof(CGL_PROF_HEAD, 8.0 * G * rows * (double)h.H, 0.0, st);   // last hidden read, its gradient written
  // the opt-in attribute is raised only when a launch needs more than any launch before it (no runtime call per launch:
  // a captured round -- MDStyleSim.round_graph -- replays exactly the kernels an eager round launched)
  // (the attribute is per device: the high-water marks are kept per device ordinal)
  static size_t attr_stream_dev[64] = {}, attr_plain_dev[64] = {};
  int dev = 0;
  CGL_CHECK_CUDA(cudaGetDevice(&dev));
  size_t& attr_stream = attr_stream_dev[dev & 63];
  size_t& attr_plain = attr_plain_dev[dev & 63];
  if (attr_plain == 0) attr_plain = 48 * 1024;
  if (stream_ok) {
    if (smem_stream > attr_stream) {
      CGL_CHECK_CUDA(cudaFuncSetAttribute(head_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_stream));
      attr_stream = smem_stream;
    }
    head_stream_kernel<<<G, HEAD_THREADS, smem_stream, st>>>(h);
  } else {
    if (smem > attr_plain) {
      CGL_CHECK_CUDA(cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_plain = smem;
    }
    head_kernel<<<G, HEAD_THREADS, smem, st>>>(h);
  }
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

// ---- K1: the fused, shared-memory-resident client step (client_fused.cuh) ----------------------------------------
// cgl_set_fused_client_step(0 / 1) (or CGL_K1=0 / 1 in the environment) switches it; the FFMA-only / tcgen05-only GEMM
// modes of cgl_set_gemm_mode keep the layered kernels, so the parity tests that run in two modes compare the two.
static int initial_k1() {
  const char* e = getenv("CGL_K1");
  return (e && (e[0] == '0' || e[0] == '1') && e[1] == 0) ? e[0] - '0' : K1_DEFAULT_ON;
}
static int g_k1_on = initial_k1();
static bool k1_enabled() { return g_k1_on != 0 && g_gemm_mode == GEMM_AUTO; }
static bool k1_eligible(const cgl_mlp_desc* a, const cgl_mlp_layout& lay, int loss_kind, int B, const void* params,
                        const void* am, const void* av, long long ldp) {
  if (!k1_enabled()) return false;
  if (a->n_layers != 3 || a->dims[0] < 1 || a->dims[0] > K1_MAXD || a->dims[1] != K1_H1 || a->dims[2] != K1_H2 ||
      a->dims[3] != 1)
    return false;
  if (a->act[0] != CGL_ACT_LRELU || a->act[1] != CGL_ACT_LRELU) return false;
  if (loss_kind != CGL_LOSS_BCE && loss_kind != CGL_LOSS_MSE) return false;
  if (B < 1 || 2 * B > K1_MAXROWS) return false;
  if (ldp % 4 != 0 || lay.w_off[1] % 4 != 0 || !aligned16(params) || (am && !aligned16(am)) || (av && !aligned16(av)))
    return false;
  return true;
}
static int launch_k1(const cgl_mlp_desc* a, const cgl_mlp_layout& lay, int G, float* params, float* am, float* av,
                     long long ldp, int* step, const int* ids, const float* real, const int* n_real, const float* fake,
                     const int* fake_idx, const float* xg, const int* xg_idx, int B, int loss_kind, float d_scale,
                     const cgl_train_cfg* cfg, float* out_dloss, float* out_gloss, float* out_dxg, int do_d, int do_g,
                     cudaStream_t st) {
  K1Params k = {};
  k.d = a->dims[0]; k.B = B;
  k.params = params; k.adam_m = am; k.adam_v = av; k.ldp = ldp; k.ids = ids; k.step = step;
  for (int l = 0; l < 3; ++l) { k.w_off[l] = lay.w_off[l]; k.b_off[l] = lay.b_off[l]; }
  k.real = real; k.n_real = n_real; k.fake = fake; k.fake_idx = fake_idx; k.xg = xg; k.xg_idx = xg_idx;
  k.out_dloss = out_dloss; k.out_gloss = out_gloss; k.out_dxg = out_dxg;
  k.loss_kind = loss_kind; k.last_act = a->act[2]; k.slope = a->lrelu_slope; k.d_scale = d_scale;
  if (cfg) { k.lr = cfg->lr; k.b1 = cfg->beta1; k.b2 = cfg->beta2; k.eps = cfg->eps; }
  k.do_d = do_d; k.do_g = do_g;
  static unsigned long long attr_set = 0;
  if (first_use_on_device(attr_set))
    CGL_CHECK_CUDA(cudaFuncSetAttribute(client_step_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)K1_SMEM_BYTES));
  // algorithmic work: MACs = B (8 M - 2 M1) for both parts, 24 B per parameter + the batches
  const double M1 = (double)a->dims[0] * K1_H1, M = M1 + (double)K1_H1 * K1_H2 + K1_H2;
  const double macs = (do_d ? 2.0 * B * (3.0 * M - M1) : 0.0) + (do_g ? 2.0 * B * M : 0.0);
  const double bytes = (do_d ? 24.0 * (double)lay.n_params + 8.0 * B * a->dims[0] : 4.0 * (double)lay.n_params) +
                       (do_g ? 8.0 * B * a->dims[0] : 0.0);
  ProfScope prof(CGL_PROF_CLIENT_FUSED, G * bytes, 2.0 * G * macs, st);
  client_step_fused_kernel<<<G, K1_THREADS, K1_SMEM_BYTES, st>>>(k);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

}  // namespace cgl

using namespace cgl;

extern "C" size_t cgl_d_step_workspace_bytes(const cgl_mlp_desc* arch, int G, int B) {
  if (!arch || G <= 0 || B <= 0) return 0;
  return ws_bytes(arch, G, 2 * B);
}
extern "C" size_t cgl_g_loss_workspace_bytes(const cgl_mlp_desc* arch, int G, int B) {
  if (!arch || G <= 0 || B <= 0) return 0;
  return ws_bytes(arch, G, B);
}

extern "C" int cgl_d_step(const cgl_mlp_desc* arch, int G, float* params, float* adam_m, float* adam_v, int64_t ldp,
                          int32_t* step, const int32_t* client_ids, const float* real, const int32_t* n_real,
                          const float* fake, const int32_t* fake_idx, int B, const cgl_train_cfg* cfg,
                          float* out_dloss, void* workspace, size_t workspace_bytes, cgl_stream_t stream) {
  CGL_REQUIRE(cfg != nullptr, "cfg is NULL");
  int rc = validate_d_arch(arch, cfg->loss_kind);
  if (rc) return rc;
  if (G == 0) return CGL_OK;
  CGL_REQUIRE(G > 0 && G <= 65535, "G=%d out of range (1..65535 groups per call)", G);
  CGL_REQUIRE(B > 0, "B must be positive");
  CGL_REQUIRE(params && adam_m && adam_v && step && real && fake && out_dloss && workspace, "NULL tensor pointer");
  cgl_mlp_layout lay;
  rc = cgl_mlp_layout_of(arch, &lay);
  if (rc) return rc;
  CGL_REQUIRE(ldp >= lay.n_params, "ldp=%lld smaller than packed row (%lld)", (long long)ldp, (long long)lay.n_params);
  const int rows = 2 * B;
  if (workspace_bytes < ws_bytes(arch, G, rows)) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, ws_bytes(arch, G, rows));
    return CGL_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (k1_eligible(arch, lay, cfg->loss_kind, B, params, adam_m, adam_v, ldp))
    return launch_k1(arch, lay, G, params, adam_m, adam_v, ldp, step, client_ids, real, n_real, fake, fake_idx, nullptr,
                     nullptr, B, cfg->loss_kind, cfg->d_loss_scale, cfg, out_dloss, nullptr, nullptr, 1, 0, st);
  Workspace w = carve(arch, G, rows, workspace);
  const int L = arch->n_layers;
  const int d = arch->dims[0];

  adam_prepare_kernel<<<(G + 127) / 128, 128, 0, st>>>(G, step, client_ids, cfg->lr, cfg->beta1, cfg->beta2, cfg->eps,
                                                      w.scal);
  CGL_CHECK_LAUNCH();

  // rows [0,B): real[g], rows [B,2B): fake[fake_idx[g]]   (CGLGAN/2DMG/main.py:361-363)
  RowMap X = dual_rows(real, (long long)B * d, nullptr, B, fake, (long long)B * d, fake_idx, d);
  rc = forward_hidden(arch, lay, G, rows, X, params, ldp, client_ids, w, st);
  if (rc) return rc;
  // loss(D(real), 1) + loss(D(fake), 0); Adam on the last layer
  rc = launch_head(arch, lay, G, rows, B, w, params, adam_m, adam_v, ldp, client_ids, step, cfg, cfg->loss_kind,
                   cfg->d_loss_scale, n_real, 1.f, 0.f, out_dloss, 1, st);
  if (rc) return rc;
  for (int l = L - 2; l >= 0; --l) {
    const int in = arch->dims[l], out = arch->dims[l + 1];
    if (l > 0) {
      CGL_CHECK_CUDA(run_linear_bwd_data(G, rows, in, out, w.dZ[l + 1], (long long)rows * out, params, ldp, client_ids,
                                         lay.w_off[l], w.H[l], (long long)rows * in, arch->act[l - 1],
                                         arch->lrelu_slope, w.dZ[l], (long long)rows * in, st));
    }
    RowMap Xin = (l == 0) ? X : single_rows(w.H[l], (long long)rows * in, nullptr, in);
    const AdamArgs ad = {adam_m, adam_v, step, cfg->lr, cfg->beta1, cfg->beta2, cfg->eps, w.scal};
    CGL_CHECK_CUDA(run_linear_wgrad(G, rows, in, out, w.dZ[l + 1], (long long)rows * out, Xin, params, ldp, client_ids,
                                    lay.w_off[l], lay.b_off[l], &ad, st));
  }
  return CGL_OK;
}

extern "C" int cgl_g_loss(const cgl_mlp_desc* arch, int G, const float* params, int64_t ldp,
                          const int32_t* client_ids, const float* xg, const int32_t* xg_idx, int B, int loss_kind,
                          float* out_loss, float* out_dxg, void* workspace, size_t workspace_bytes,
                          cgl_stream_t stream) {
  int rc = validate_d_arch(arch, loss_kind);
  if (rc) return rc;
  if (G == 0) return CGL_OK;
  CGL_REQUIRE(G > 0 && G <= 65535, "G=%d out of range (1..65535 groups per call)", G);
  CGL_REQUIRE(B > 0, "B must be positive");
  CGL_REQUIRE(params && xg && out_loss && workspace, "NULL tensor pointer");
  cgl_mlp_layout lay;
  rc = cgl_mlp_layout_of(arch, &lay);
  if (rc) return rc;
  CGL_REQUIRE(ldp >= lay.n_params, "ldp smaller than packed row");
  if (workspace_bytes < ws_bytes(arch, G, B)) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, ws_bytes(arch, G, B));
    return CGL_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (k1_eligible(arch, lay, loss_kind, B, params, nullptr, nullptr, ldp))
    return launch_k1(arch, lay, G, const_cast<float*>(params), nullptr, nullptr, ldp, nullptr, client_ids, nullptr, nullptr,
                     nullptr, nullptr, xg, xg_idx, B, loss_kind, 1.f, nullptr, nullptr, out_loss, out_dxg, 0, 1, st);
  Workspace w = carve(arch, G, B, workspace);
  const int L = arch->n_layers;
  const int d = arch->dims[0];
  RowMap X = single_rows(xg, (long long)B * d, xg_idx, d);
  rc = forward_hidden(arch, lay, G, B, X, params, ldp, client_ids, w, st);
  if (rc) return rc;
  // G_loss = loss(D(Xg), valid)   (CGLGAN/2DMG/main.py:368-372); no parameter update
  rc = launch_head(arch, lay, G, B, B, w, const_cast<float*>(params), nullptr, nullptr, ldp, client_ids, nullptr,
                   nullptr, loss_kind, 1.f, nullptr, 1.f, 1.f, out_loss, 0, st);
  if (rc) return rc;
  if (!out_dxg) return CGL_OK;
  for (int l = L - 2; l >= 0; --l) {
    const int in = arch->dims[l], out = arch->dims[l + 1];
    if (l > 0) {
      CGL_CHECK_CUDA(run_linear_bwd_data(G, B, in, out, w.dZ[l + 1], (long long)B * out, params, ldp, client_ids,
                                         lay.w_off[l], w.H[l], (long long)B * in, arch->act[l - 1], arch->lrelu_slope,
                                         w.dZ[l], (long long)B * in, st));
    } else {
      CGL_CHECK_CUDA(run_linear_bwd_data(G, B, in, out, w.dZ[1], (long long)B * out, params, ldp, client_ids,
                                         lay.w_off[0], nullptr, 0, CGL_ACT_NONE, 0.f, out_dxg, (long long)B * in, st));
    }
  }
  return CGL_OK;
}

// The whole client step of a round in one call: the D step on (real, fake), then the G loss and dLoss/dXg through the
// UPDATED discriminator (Worker.train, CGLGAN/2DMG/main.py:344-375 with epoch == 1). One launch per call where the fused
// shared-memory-resident kernel applies (client_fused.cuh), otherwise cgl_d_step followed by cgl_g_loss.
extern "C" int cgl_client_step(const cgl_mlp_desc* arch, int G, float* params, float* adam_m, float* adam_v, int64_t ldp,
                               int32_t* step, const int32_t* client_ids, const float* real, const int32_t* n_real,
                               const float* fake, const int32_t* fake_idx, const float* xg, const int32_t* xg_idx, int B,
                               const cgl_train_cfg* cfg, float* out_dloss, float* out_gloss, float* out_dxg,
                               void* workspace, size_t workspace_bytes, cgl_stream_t stream) {
  CGL_REQUIRE(cfg != nullptr, "cfg is NULL");
  int rc = validate_d_arch(arch, cfg->loss_kind);
  if (rc) return rc;
  if (G == 0) return CGL_OK;
  CGL_REQUIRE(G > 0 && G <= 65535, "G=%d out of range (1..65535 groups per call)", G);
  CGL_REQUIRE(B > 0, "B must be positive");
  CGL_REQUIRE(params && adam_m && adam_v && step && real && fake && xg && out_dloss && out_gloss, "NULL tensor pointer");
  cgl_mlp_layout lay;
  rc = cgl_mlp_layout_of(arch, &lay);
  if (rc) return rc;
  CGL_REQUIRE(ldp >= lay.n_params, "ldp=%lld smaller than packed row (%lld)", (long long)ldp, (long long)lay.n_params);
  if (k1_eligible(arch, lay, cfg->loss_kind, B, params, adam_m, adam_v, ldp))
    return launch_k1(arch, lay, G, params, adam_m, adam_v, ldp, step, client_ids, real, n_real, fake, fake_idx, xg, xg_idx,
                     B, cfg->loss_kind, cfg->d_loss_scale, cfg, out_dloss, out_gloss, out_dxg, 1, 1,
                     (cudaStream_t)stream);
  rc = cgl_d_step(arch, G, params, adam_m, adam_v, ldp, step, client_ids, real, n_real, fake, fake_idx, B, cfg, out_dloss,
                  workspace, workspace_bytes, stream);
  if (rc) return rc;
  return cgl_g_loss(arch, G, params, ldp, client_ids, xg, xg_idx, B, cfg->loss_kind, out_gloss, out_dxg, workspace,
                    workspace_bytes, stream);
}

extern "C" int cgl_set_gemm_mode(int mode) {
  CGL_REQUIRE(mode == GEMM_AUTO || mode == GEMM_FFMA || mode == GEMM_TC, "gemm mode must be 0 (auto), 1 (FFMA) or 2 (tcgen05)");
  g_gemm_mode = mode;
  return CGL_OK;
}
extern "C" int cgl_get_gemm_mode(void) { return g_gemm_mode; }
extern "C" int cgl_set_fused_client_step(int on) {
  g_k1_on = on ? 1 : 0;
  return CGL_OK;
}
extern "C" int cgl_get_fused_client_step(void) { return g_k1_on; }

// Bring-up only: CTA (0,0,g) of every tcgen05 GEMM launched from this translation unit (cgl_d_step, cgl_g_loss,
// cgl_linear_*) stamps clock64() at its milestones into buf[g*16 + i] (csrc/tc_gemm.cuh). NULL switches it off.
namespace cgl { cudaError_t set_timeline_gstep(long long* device_buf); }  // the other module's copy of the symbol
extern "C" int cgl_debug_set_timeline(long long* device_buf) {
  CGL_CHECK_CUDA(cudaMemcpyToSymbol(g_tc_timeline, &device_buf, sizeof(device_buf)));
  CGL_CHECK_CUDA(set_timeline_gstep(device_buf));
  return CGL_OK;
}

// ---- building blocks ------------------------------------------------------------------------
extern "C" int cgl_linear_fwd(int G, int rows, int in, int out, const float* x, int64_t x_gstride, const float* wbase,
                              int64_t ldp, const int32_t* ids, int64_t w_off, int64_t b_off, int act, float slope,
                              float* y, int64_t y_gstride, cgl_stream_t stream) {
  CGL_REQUIRE(G >= 0 && G <= 65535 && rows > 0 && in > 0 && out > 0, "bad shape");
  CGL_REQUIRE(x && wbase && y, "NULL tensor pointer");
  RowMap X = single_rows(x, x_gstride, nullptr, in);
  CGL_CHECK_CUDA(run_linear_fwd(G, rows, in, out, X, wbase, ldp, ids, w_off, b_off, act, slope, y, y_gstride,
                                (cudaStream_t)stream));
  return CGL_OK;
}

extern "C" int cgl_linear_bwd_data(int G, int rows, int in, int out, const float* dy, int64_t dy_gstride,
                                   const float* wbase, int64_t ldp, const int32_t* ids, int64_t w_off,
                                   const float* saved, int64_t saved_gstride, int act, float slope, float* dx,
                                   int64_t dx_gstride, cgl_stream_t stream) {
  CGL_REQUIRE(G >= 0 && G <= 65535 && rows > 0 && in > 0 && out > 0, "bad shape");
  CGL_REQUIRE(dy && wbase && dx, "NULL tensor pointer");
  CGL_CHECK_CUDA(run_linear_bwd_data(G, rows, in, out, dy, dy_gstride, wbase, ldp, ids, w_off, saved, saved_gstride, act,
                                     slope, dx, dx_gstride, (cudaStream_t)stream));
  return CGL_OK;
}

extern "C" int cgl_linear_wgrad(int G, int rows, int in, int out, const float* dy, int64_t dy_gstride, const float* x,
                                int64_t x_gstride, float* gbase, int64_t ldg, const int32_t* ids, int64_t w_off,
                                int64_t b_off, cgl_stream_t stream) {
  CGL_REQUIRE(G >= 0 && G <= 65535 && rows > 0 && in > 0 && out > 0, "bad shape");
  CGL_REQUIRE(dy && x && gbase, "NULL tensor pointer");
  RowMap X = single_rows(x, x_gstride, nullptr, in);
  CGL_CHECK_CUDA(run_linear_wgrad(G, rows, in, out, dy, dy_gstride, X, gbase, ldg, ids, w_off, b_off, nullptr,
                                  (cudaStream_t)stream));
  return CGL_OK;
}

extern "C" int cgl_linear_wgrad_adam(int G, int rows, int in, int out, const float* dy, int64_t dy_gstride, const float* x,
                                     int64_t x_gstride, float* params, float* adam_m, float* adam_v, int64_t ld,
                                     const int32_t* step, const int32_t* ids, int64_t w_off, int64_t b_off, float lr,
                                     float beta1, float beta2, float eps, void* adam_scratch, cgl_stream_t stream) {
  CGL_REQUIRE(G >= 0 && G <= 65535 && rows > 0 && in > 0 && out > 0, "bad shape");
  CGL_REQUIRE(dy && x && params && adam_m && adam_v && step, "NULL tensor pointer");
  CGL_REQUIRE(b_off >= 0, "the fused Adam epilogue updates weight and bias together");
  RowMap X = single_rows(x, x_gstride, nullptr, in);
  AdamArgs ad = {adam_m, adam_v, step, lr, beta1, beta2, eps, nullptr};
  if (adam_scratch && G > 0) {   // the step's scalars once per group (what cgl_d_step / cgl_mlp_backward do)
    adam_scalars_kernel<<<(G + 127) / 128, 128, 0, (cudaStream_t)stream>>>(G, step, ids, lr, beta1, beta2, eps,
                                                                          (AdamScalars*)adam_scratch);
    CGL_CHECK_LAUNCH();
    ad.scal = (const AdamScalars*)adam_scratch;
  }
  CGL_CHECK_CUDA(run_linear_wgrad(G, rows, in, out, dy, dy_gstride, X, params, ld, ids, w_off, b_off, &ad,
                                  (cudaStream_t)stream));
  return CGL_OK;
}
