// gstep.cu -- the generator side of a round (row a4 / a5 of SURVEY.md section 8): forward of a stack of
// Linear [+ BatchNorm1d(eps = 0.8, batch statistics)] + LeakyReLU / Tanh over G independent groups
// (edge servers, heads, or FL clients), and its backward with the Adam step fused in.
// Reference: Server.train, CGLGAN/2DMG/main.py:229-234,254-276; mixed-gan.py:238-292;
//            model/mnist_model.py:5-29 (Generator), :32-66 (MixGenerator); FL: FLGAN/MNIST/flgan.py:251-269.
#include "linear.cuh"

namespace cgl {

// ---------------------------------------------------------------------------------------------
// BatchNorm1d, training mode, one thread per (group, feature); rows are walked three times (the
// [rows x 128] slab of a CTA stays in L1). torch semantics: biased variance for the normalisation,
// unbiased for running_var, momentum 0.1 (model/mnist_model.py:13 passes eps = 0.8 positionally).
//   u [G][rows][F] -> h = act((u - mean) * invstd * gamma + beta);  saves mean / invstd per (g, f)
// ---------------------------------------------------------------------------------------------
struct BnFwdParams {
  int rows, F;
  const float* u; long long u_gstride;
  float* h; long long h_gstride;
  const float* params; long long ldp; const int* ids; long long gamma_off, beta_off;
  float* stats; long long ld_stats; long long mean_off, var_off;  // running stats row (NULL: none)
  float* save_mean; float* save_invstd;                            // [G][F]
  float eps, momentum; int train; int act; float slope;
};

__global__ void __launch_bounds__(128) bn_fwd_kernel(const BnFwdParams p) {
  const int g = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= p.F) return;
  const int rowid = p.ids ? p.ids[g] : g;
  const float* u = p.u + (long long)g * p.u_gstride + f;
  float* h = p.h + (long long)g * p.h_gstride + f;
  const float gamma = __ldg(p.params + (long long)rowid * p.ldp + p.gamma_off + f);
  const float beta = __ldg(p.params + (long long)rowid * p.ldp + p.beta_off + f);
  float mean, var;
  const int n = p.rows;
  if (p.train) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int r = 0;
    for (; r + 3 < n; r += 4) {
      s0 += u[(long long)r * p.F]; s1 += u[(long long)(r + 1) * p.F];
      s2 += u[(long long)(r + 2) * p.F]; s3 += u[(long long)(r + 3) * p.F];
    }
    for (; r < n; ++r) s0 += u[(long long)r * p.F];
    mean = ((s0 + s1) + (s2 + s3)) / (float)n;
    s0 = s1 = s2 = s3 = 0.f;
    r = 0;
    for (; r + 3 < n; r += 4) {
      float d0 = u[(long long)r * p.F] - mean, d1 = u[(long long)(r + 1) * p.F] - mean;
      float d2 = u[(long long)(r + 2) * p.F] - mean, d3 = u[(long long)(r + 3) * p.F] - mean;
      s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s2 = fmaf(d2, d2, s2); s3 = fmaf(d3, d3, s3);
    }
    for (; r < n; ++r) { float d0 = u[(long long)r * p.F] - mean; s0 = fmaf(d0, d0, s0); }
    const float ss = (s0 + s1) + (s2 + s3);
    var = ss / (float)n;
    if (p.stats) {
      float* rm = p.stats + (long long)rowid * p.ld_stats + p.mean_off + f;
      float* rv = p.stats + (long long)rowid * p.ld_stats + p.var_off + f;
      const float unbiased = n > 1 ? ss / (float)(n - 1) : var;
      *rm = (1.f - p.momentum) * *rm + p.momentum * mean;
      *rv = (1.f - p.momentum) * *rv + p.momentum * unbiased;
    }
  } else {
    mean = p.stats[(long long)rowid * p.ld_stats + p.mean_off + f];
    var = p.stats[(long long)rowid * p.ld_stats + p.var_off + f];
  }
  const float invstd = 1.f / sqrtf(var + p.eps);
  if (p.save_mean) {
    p.save_mean[(long long)g * p.F + f] = mean;
    p.save_invstd[(long long)g * p.F + f] = invstd;
  }
  const float a = invstd * gamma;
  for (int r = 0; r < n; ++r) {
    const float y = (u[(long long)r * p.F] - mean) * a + beta;
    h[(long long)r * p.F] = act_fwd(y, p.act, p.slope);
  }
}

// The same kernel with its 128-feature column block staged in shared memory first: every global load of the
// block is in flight at once (16-byte cp.async, one 512-byte row segment per warp request) instead of three
// dependent passes of 4-byte loads per thread; the arithmetic (and its order) is unchanged, so the results are
// bit-identical. Needs rows * 512 bytes of shared memory and float4-addressable rows.
__device__ __forceinline__ void bn_cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void bn_cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
// tile[r][0..127] <- src[r * F + f0 .. f0 + 127] (zero beyond F), all 128 threads
__device__ __forceinline__ void bn_stage_block(float* tile, const float* src, int rows, int F, int f0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int f = f0 + lane * 4;
  for (int r = warp; r < rows; r += 4) {
    if (f < F) bn_cp_async16(tile + r * 128 + lane * 4, src + (long long)r * F + f);
    else *reinterpret_cast<float4*>(tile + r * 128 + lane * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(128) bn_fwd_smem_kernel(const BnFwdParams p) {
  extern __shared__ __align__(16) float bn_tile[];   // [rows][128]
  const int g = blockIdx.y;
  const int f0 = blockIdx.x * 128;
  const int t = threadIdx.x;
  const int f = f0 + t;
  const int n = p.rows;
  bn_stage_block(bn_tile, p.u + (long long)g * p.u_gstride, n, p.F, f0);
  bn_cp_async_wait_all();
  __syncthreads();
  if (f >= p.F) return;
  const int rowid = p.ids ? p.ids[g] : g;
  const float* u = bn_tile + t;                      // this thread's column, stride 128
  float* h = p.h + (long long)g * p.h_gstride + f;
  const float gamma = __ldg(p.params + (long long)rowid * p.ldp + p.gamma_off + f);
  const float beta = __ldg(p.params + (long long)rowid * p.ldp + p.beta_off + f);
  float mean, var;
  if (p.train) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int r = 0;
    for (; r + 3 < n; r += 4) {
      s0 += u[r * 128]; s1 += u[(r + 1) * 128];
      s2 += u[(r + 2) * 128]; s3 += u[(r + 3) * 128];
    }
    for (; r < n; ++r) s0 += u[r * 128];
    mean = ((s0 + s1) + (s2 + s3)) / (float)n;
    s0 = s1 = s2 = s3 = 0.f;
    r = 0;
    for (; r + 3 < n; r += 4) {
      float d0 = u[r * 128] - mean, d1 = u[(r + 1) * 128] - mean;
      float d2 = u[(r + 2) * 128] - mean, d3 = u[(r + 3) * 128] - mean;
      s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s2 = fmaf(d2, d2, s2); s3 = fmaf(d3, d3, s3);
    }
    for (; r < n; ++r) { float d0 = u[r * 128] - mean; s0 = fmaf(d0, d0, s0); }
    const float ss = (s0 + s1) + (s2 + s3);
    var = ss / (float)n;
    if (p.stats) {
      float* rm = p.stats + (long long)rowid * p.ld_stats + p.mean_off + f;
      float* rv = p.stats + (long long)rowid * p.ld_stats + p.var_off + f;
      const float unbiased = n > 1 ? ss / (float)(n - 1) : var;
      *rm = (1.f - p.momentum) * *rm + p.momentum * mean;
      *rv = (1.f - p.momentum) * *rv + p.momentum * unbiased;
    }
  } else {
    mean = p.stats[(long long)rowid * p.ld_stats + p.mean_off + f];
    var = p.stats[(long long)rowid * p.ld_stats + p.var_off + f];
  }
  const float invstd = 1.f / sqrtf(var + p.eps);
  if (p.save_mean) {
    p.save_mean[(long long)g * p.F + f] = mean;
    p.save_invstd[(long long)g * p.F + f] = invstd;
  }
  const float a = invstd * gamma;
  for (int r = 0; r < n; ++r) {
    const float y = (u[r * 128] - mean) * a + beta;
    h[(long long)r * p.F] = act_fwd(y, p.act, p.slope);
  }
}

// Backward of the same: dz is the gradient wrt the BatchNorm OUTPUT (activation derivative already
// applied); overwritten in place with the gradient wrt the BatchNorm input u. gamma / beta take
// their Adam step here (torch.optim.Adam on bn.weight / bn.bias).
struct BnBwdParams {
  int rows, F;
  float* dz; long long dz_gstride;
  const float* u; long long u_gstride;
  const float* save_mean; const float* save_invstd;
  float* params; float* adam_m; float* adam_v; long long ldp; const int* ids; long long gamma_off, beta_off;
  const int* step; float lr, b1, b2, eps;
  const AdamScalars* scal;
};

__global__ void __launch_bounds__(128) bn_bwd_kernel(const BnBwdParams p) {
  const int g = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= p.F) return;
  const int rowid = p.ids ? p.ids[g] : g;
  float* dz = p.dz + (long long)g * p.dz_gstride + f;
  const float* u = p.u + (long long)g * p.u_gstride + f;
  const float mean = p.save_mean[(long long)g * p.F + f];
  const float invstd = p.save_invstd[(long long)g * p.F + f];
  const int n = p.rows;
  float sb0 = 0.f, sb1 = 0.f, sg0 = 0.f, sg1 = 0.f;
  int r = 0;
  for (; r + 1 < n; r += 2) {
    const float d0 = dz[(long long)r * p.F], d1 = dz[(long long)(r + 1) * p.F];
    sb0 += d0; sb1 += d1;
    sg0 = fmaf(d0, u[(long long)r * p.F] - mean, sg0);
    sg1 = fmaf(d1, u[(long long)(r + 1) * p.F] - mean, sg1);
  }
  for (; r < n; ++r) {
    const float d0 = dz[(long long)r * p.F];
    sb0 += d0;
    sg0 = fmaf(d0, u[(long long)r * p.F] - mean, sg0);
  }
  const float dbeta = sb0 + sb1;
  const float dotp = sg0 + sg1;          // sum dz * (u - mean)
  const float dgamma = dotp * invstd;
  const long long go = (long long)rowid * p.ldp + p.gamma_off + f;
  const long long bo = (long long)rowid * p.ldp + p.beta_off + f;
  const float gamma = p.params[go];
  // du = (dz - dbeta/n - (u - mean) * invstd^2 * dotp / n) * invstd * gamma     (ATen batch_norm_backward)
  const float k = dotp * invstd * invstd / (float)n;
  const float mb = dbeta / (float)n;
  const float a = invstd * gamma;
  for (r = 0; r < n; ++r) {
    const float d = dz[(long long)r * p.F];
    dz[(long long)r * p.F] = (d - mb - (u[(long long)r * p.F] - mean) * k) * a;
  }
  const AdamScalars s = p.scal ? p.scal[g] : make_adam_scalars(p.step[rowid], p.lr, p.b1, p.b2, p.eps);
  {
    float w = gamma, mm = p.adam_m[go], vv = p.adam_v[go];
    adam_update(w, mm, vv, dgamma, s);
    p.params[go] = w; p.adam_m[go] = mm; p.adam_v[go] = vv;
  }
  {
    float w = p.params[bo], mm = p.adam_m[bo], vv = p.adam_v[bo];
    adam_update(w, mm, vv, dbeta, s);
    p.params[bo] = w; p.adam_m[bo] = mm; p.adam_v[bo] = vv;
  }
}

// bn_bwd_kernel with dz and u staged in shared memory (see bn_fwd_smem_kernel): 2 * rows * 512 bytes, same arithmetic.
__global__ void __launch_bounds__(128) bn_bwd_smem_kernel(const BnBwdParams p) {
  extern __shared__ __align__(16) float bn_tile[];   // dz [rows][128] | u [rows][128]
  const int g = blockIdx.y;
  const int f0 = blockIdx.x * 128;
  const int t = threadIdx.x;
  const int f = f0 + t;
  const int n = p.rows;
  float* tz = bn_tile;
  float* tu = bn_tile + n * 128;
  bn_stage_block(tz, p.dz + (long long)g * p.dz_gstride, n, p.F, f0);
  bn_stage_block(tu, p.u + (long long)g * p.u_gstride, n, p.F, f0);
  bn_cp_async_wait_all();
  __syncthreads();
  if (f >= p.F) return;
  const int rowid = p.ids ? p.ids[g] : g;
  float* dzg = p.dz + (long long)g * p.dz_gstride + f;
  const float* dz = tz + t;
  const float* u = tu + t;
  const float mean = p.save_mean[(long long)g * p.F + f];
  const float invstd = p.save_invstd[(long long)g * p.F + f];
  float sb0 = 0.f, sb1 = 0.f, sg0 = 0.f, sg1 = 0.f;
  int r = 0;
  for (; r + 1 < n; r += 2) {
    const float d0 = dz[r * 128], d1 = dz[(r + 1) * 128];
    sb0 += d0; sb1 += d1;
    sg0 = fmaf(d0, u[r * 128] - mean, sg0);
    sg1 = fmaf(d1, u[(r + 1) * 128] - mean, sg1);
  }
  for (; r < n; ++r) {
    const float d0 = dz[r * 128];
    sb0 += d0;
    sg0 = fmaf(d0, u[r * 128] - mean, sg0);
  }
  const float dbeta = sb0 + sb1;
  const float dotp = sg0 + sg1;
  const float dgamma = dotp * invstd;
  const long long go = (long long)rowid * p.ldp + p.gamma_off + f;
  const long long bo = (long long)rowid * p.ldp + p.beta_off + f;
  const float gamma = p.params[go];
  const float k = dotp * invstd * invstd / (float)n;
  const float mb = dbeta / (float)n;
  const float a = invstd * gamma;
  for (r = 0; r < n; ++r) dzg[(long long)r * p.F] = (dz[r * 128] - mb - (u[r * 128] - mean) * k) * a;
  const AdamScalars s = p.scal ? p.scal[g] : make_adam_scalars(p.step[rowid], p.lr, p.b1, p.b2, p.eps);
  {
    float w = gamma, mm = p.adam_m[go], vv = p.adam_v[go];
    adam_update(w, mm, vv, dgamma, s);
    p.params[go] = w; p.adam_m[go] = mm; p.adam_v[go] = vv;
  }
  {
    float w = p.params[bo], mm = p.adam_m[bo], vv = p.adam_v[bo];
    adam_update(w, mm, vv, dbeta, s);
    p.params[bo] = w; p.adam_m[bo] = mm; p.adam_v[bo] = vv;
  }
}

// shared-memory variants apply when the block fits and the rows are float4-addressable
static inline bool bn_smem_ok(int rows, int F, const void* a, long long a_gs, const void* b, long long b_gs, int arrays) {
  return (size_t)arrays * rows * 512 <= 200 * 1024 && F % 4 == 0 && aligned16(a) && a_gs % 4 == 0 &&
         (!b || (aligned16(b) && b_gs % 4 == 0));
}

// dz = dy * act'(y), y the saved activation output (the Tanh of the generator's last layer)
__global__ void act_bwd_kernel(long long n, const float* __restrict__ dy, const float* __restrict__ y, float* dz,
                               int act, float slope) {
  long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 a = *reinterpret_cast<const float4*>(dy + i);
    const float4 b = *reinterpret_cast<const float4*>(y + i);
    float4 o;
    o.x = a.x * act_bwd_from_out(b.x, act, slope); o.y = a.y * act_bwd_from_out(b.y, act, slope);
    o.z = a.z * act_bwd_from_out(b.z, act, slope); o.w = a.w * act_bwd_from_out(b.w, act, slope);
    *reinterpret_cast<float4*>(dz + i) = o;
  } else {
    for (; i < n; ++i) dz[i] = dy[i] * act_bwd_from_out(y[i], act, slope);
  }
}

// ---- workspace ----------------------------------------------------------------------------------
static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

struct MlpWs {
  float* H[CGL_MAX_LAYERS + 1];   // H[l], l = 1..L-1: input of layer l (post-activation output of layer l-1)
  float* U[CGL_MAX_LAYERS];       // pre-BatchNorm linear output of layer l (bn[l] only)
  float* mean[CGL_MAX_LAYERS];
  float* invstd[CGL_MAX_LAYERS];
  float* dZ[2];                   // backward ping-pong, G*rows*maxdim each
  AdamScalars* scal;              // [G] scalars of the current Adam step
  size_t bytes;
};
static MlpWs mlp_carve(const cgl_mlp_desc* a, int G, int rows, void* base) {
  MlpWs w = {};
  char* p = (char*)base;
  size_t off = 0;
  int maxdim = 0;
  for (int l = 0; l <= a->n_layers; ++l) maxdim = a->dims[l] > maxdim ? a->dims[l] : maxdim;
  for (int l = 0; l < a->n_layers; ++l) {
    if (l > 0) { w.H[l] = (float*)(p + off); off += up256((size_t)G * rows * a->dims[l] * 4); }
    if (a->bn[l]) {
      w.U[l] = (float*)(p + off); off += up256((size_t)G * rows * a->dims[l + 1] * 4);
      w.mean[l] = (float*)(p + off); off += up256((size_t)G * a->dims[l + 1] * 4);
      w.invstd[l] = (float*)(p + off); off += up256((size_t)G * a->dims[l + 1] * 4);
    }
  }
  for (int i = 0; i < 2; ++i) { w.dZ[i] = (float*)(p + off); off += up256((size_t)G * rows * maxdim * 4); }
  w.scal = (AdamScalars*)(p + off); off += up256((size_t)G * sizeof(AdamScalars));
  w.bytes = off + 256;
  return w;
}

static int validate_mlp(const cgl_mlp_desc* a) {
  CGL_REQUIRE(a != nullptr, "arch is NULL");
  CGL_REQUIRE(a->n_layers >= 1 && a->n_layers <= CGL_MAX_LAYERS, "n_layers=%d out of range", a->n_layers);
  for (int i = 0; i < a->n_layers; ++i) CGL_REQUIRE(a->dims[i] > 0 && a->dims[i + 1] > 0, "bad width at layer %d", i);
  return CGL_OK;
}

cudaError_t set_timeline_gstep(long long* device_buf) {
  return cudaMemcpyToSymbol(g_tc_timeline, &device_buf, sizeof(device_buf));
}

}  // namespace cgl

using namespace cgl;

extern "C" size_t cgl_mlp_workspace_bytes(const cgl_mlp_desc* arch, int G, int rows) {
  if (!arch || G <= 0 || rows <= 0 || arch->n_layers < 1 || arch->n_layers > CGL_MAX_LAYERS) return 0;
  return mlp_carve(arch, G, rows, nullptr).bytes;
}

extern "C" int cgl_mlp_forward(const cgl_mlp_desc* arch, int G, const float* params, int64_t ldp, const int32_t* ids,
                               float* bn_stats, int64_t ld_stats, int train, const float* x, int64_t x_gstride,
                               const int32_t* x_idx, int rows, float* y, void* workspace, size_t workspace_bytes,
                               cgl_stream_t stream) {
  int rc = validate_mlp(arch);
  if (rc) return rc;
  if (G == 0) return CGL_OK;
  CGL_REQUIRE(G > 0 && G <= 65535, "G=%d out of range (1..65535 groups per call)", G);
  CGL_REQUIRE(rows > 0, "rows must be positive");
  CGL_REQUIRE(params && x && y && workspace, "NULL tensor pointer");
  cgl_mlp_layout lay;
  rc = cgl_mlp_layout_of(arch, &lay);
  if (rc) return rc;
  CGL_REQUIRE(ldp >= lay.n_params, "ldp=%lld smaller than packed row (%lld)", (long long)ldp, (long long)lay.n_params);
  CGL_REQUIRE(lay.n_bn_stats == 0 || (bn_stats && ld_stats >= lay.n_bn_stats), "BatchNorm layers need a running-stats buffer");
  MlpWs w = mlp_carve(arch, G, rows, workspace);
  if (workspace_bytes < w.bytes) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, w.bytes);
    return CGL_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int L = arch->n_layers;
  for (int l = 0; l < L; ++l) {
    const int in = arch->dims[l], out = arch->dims[l + 1];
    RowMap X = (l == 0) ? single_rows(x, x_gstride, x_idx, in) : single_rows(w.H[l], (long long)rows * in, nullptr, in);
    float* dst = (l + 1 < L) ? w.H[l + 1] : y;
    if (!arch->bn[l]) {
      CGL_CHECK_CUDA(run_linear_fwd(G, rows, in, out, X, params, ldp, ids, lay.w_off[l], lay.b_off[l], arch->act[l],
                                    arch->lrelu_slope, dst, (long long)rows * out, st));
    } else {
      CGL_CHECK_CUDA(run_linear_fwd(G, rows, in, out, X, params, ldp, ids, lay.w_off[l], lay.b_off[l], CGL_ACT_NONE,
                                    0.f, w.U[l], (long long)rows * out, st));
      BnFwdParams b = {};
      b.rows = rows; b.F = out;
      b.u = w.U[l]; b.u_gstride = (long long)rows * out;
      b.h = dst; b.h_gstride = (long long)rows * out;
      b.params = params; b.ldp = ldp; b.ids = ids; b.gamma_off = lay.bn_w_off[l]; b.beta_off = lay.bn_b_off[l];
      b.stats = bn_stats; b.ld_stats = ld_stats; b.mean_off = lay.bn_mean_off[l]; b.var_off = lay.bn_var_off[l];
      b.save_mean = w.mean[l]; b.save_invstd = w.invstd[l];
      b.eps = arch->bn_eps; b.momentum = arch->bn_momentum; b.train = train;
      b.act = arch->act[l]; b.slope = arch->lrelu_slope;
      dim3 grid((out + 127) / 128, G);
      ProfScope prof(CGL_PROF_BN_FWD, 8.0 * G * rows * (double)out, 0.0, st);   // u read, h written
      if (bn_smem_ok(rows, out, b.u, b.u_gstride, nullptr, 0, 1)) {
        static unsigned long long attr = 0;
        if (first_use_on_device(attr))
          CGL_CHECK_CUDA(cudaFuncSetAttribute(bn_fwd_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        bn_fwd_smem_kernel<<<grid, 128, (size_t)rows * 512, st>>>(b);
      } else {
        bn_fwd_kernel<<<grid, 128, 0, st>>>(b);
      }
      CGL_CHECK_LAUNCH();
    }
  }
  return CGL_OK;
}

extern "C" int cgl_mlp_backward(const cgl_mlp_desc* arch, int G, float* params, float* adam_m, float* adam_v,
                                int64_t ldp, int32_t* step, const int32_t* ids, const cgl_train_cfg* cfg,
                                const float* x, int64_t x_gstride, const int32_t* x_idx, int rows, const float* y,
                                const float* dy, float* dx, void* workspace, size_t workspace_bytes,
                                cgl_stream_t stream) {
  int rc = validate_mlp(arch);
  if (rc) return rc;
  if (G == 0) return CGL_OK;
  CGL_REQUIRE(cfg != nullptr, "cfg is NULL");
  CGL_REQUIRE(G > 0 && G <= 65535, "G=%d out of range (1..65535 groups per call)", G);
  CGL_REQUIRE(rows > 0, "rows must be positive");
  CGL_REQUIRE(params && adam_m && adam_v && step && x && y && dy && workspace, "NULL tensor pointer");
  cgl_mlp_layout lay;
  rc = cgl_mlp_layout_of(arch, &lay);
  if (rc) return rc;
  CGL_REQUIRE(ldp >= lay.n_params, "ldp smaller than packed row");
  MlpWs w = mlp_carve(arch, G, rows, workspace);
  if (workspace_bytes < w.bytes) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, w.bytes);
    return CGL_EWORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int L = arch->n_layers;
  adam_prepare_kernel<<<(G + 127) / 128, 128, 0, st>>>(G, step, ids, cfg->lr, cfg->beta1, cfg->beta2, cfg->eps, w.scal);
  CGL_CHECK_LAUNCH();

  // gradient wrt the last layer's pre-activation
  int cur = 0;
  {
    const long long n = (long long)G * rows * arch->dims[L];
    const long long nthreads = (n + 3) / 4;
    ProfScope prof(CGL_PROF_ELEMENTWISE, 12.0 * (double)n, 0.0, st);
    act_bwd_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, st>>>(n, dy, y, w.dZ[cur], arch->act[L - 1],
                                                                       arch->lrelu_slope);
    CGL_CHECK_LAUNCH();
  }
  const AdamArgs ad = {adam_m, adam_v, step, cfg->lr, cfg->beta1, cfg->beta2, cfg->eps, w.scal};
  for (int l = L - 1; l >= 0; --l) {
    const int in = arch->dims[l], out = arch->dims[l + 1];
    float* dU = w.dZ[cur];
    if (arch->bn[l]) {
      BnBwdParams b = {};
      b.rows = rows; b.F = out;
      b.dz = dU; b.dz_gstride = (long long)rows * out;
      b.u = w.U[l]; b.u_gstride = (long long)rows * out;
      b.save_mean = w.mean[l]; b.save_invstd = w.invstd[l];
      b.params = params; b.adam_m = adam_m; b.adam_v = adam_v; b.ldp = ldp; b.ids = ids;
      b.gamma_off = lay.bn_w_off[l]; b.beta_off = lay.bn_b_off[l];
      b.step = step; b.lr = cfg->lr; b.b1 = cfg->beta1; b.b2 = cfg->beta2; b.eps = cfg->eps; b.scal = w.scal;
      dim3 grid((out + 127) / 128, G);
      ProfScope prof(CGL_PROF_BN_BWD, 12.0 * G * rows * (double)out, 0.0, st);  // dz, u read, du written
      if (bn_smem_ok(rows, out, b.dz, b.dz_gstride, b.u, b.u_gstride, 2)) {
        static unsigned long long attr = 0;
        if (first_use_on_device(attr))
          CGL_CHECK_CUDA(cudaFuncSetAttribute(bn_bwd_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        bn_bwd_smem_kernel<<<grid, 128, (size_t)rows * 1024, st>>>(b);
      } else {
        bn_bwd_kernel<<<grid, 128, 0, st>>>(b);
      }
      CGL_CHECK_LAUNCH();
    }
    // data gradient first: it reads W_l, which the weight-gradient kernel below overwrites (fused Adam)
    if (l > 0) {
      CGL_CHECK_CUDA(run_linear_bwd_data(G, rows, in, out, dU, (long long)rows * out, params, ldp, ids, lay.w_off[l],
                                         w.H[l], (long long)rows * in, arch->act[l - 1], arch->lrelu_slope,
                                         w.dZ[cur ^ 1], (long long)rows * in, st));
    } else if (dx) {
      CGL_CHECK_CUDA(run_linear_bwd_data(G, rows, in, out, dU, (long long)rows * out, params, ldp, ids, lay.w_off[0],
                                         nullptr, 0, CGL_ACT_NONE, 0.f, dx, (long long)rows * in, st));
    }
    RowMap Xin = (l == 0) ? single_rows(x, x_gstride, x_idx, in) : single_rows(w.H[l], (long long)rows * in, nullptr, in);
    CGL_CHECK_CUDA(run_linear_wgrad(G, rows, in, out, dU, (long long)rows * out, Xin, params, ldp, ids, lay.w_off[l],
                                    lay.b_off[l], &ad, st));
    cur ^= 1;
  }
  return CGL_OK;
}

// ---- building blocks (the convolutional networks of conv.cu compose them; csrc/conv.cu, cgl-gan_b200/conv.py) ---------
// BatchNorm over the rows of [G][rows][F] with its affine parameters in packed rows: BatchNorm1d over a batch, or
// BatchNorm2d when the rows are the pixels of NHWC images (statistics over N*H*W per channel).
extern "C" int cgl_bn_forward(int G, int rows, int F, const float* u, float* h, const float* params, int64_t ldp,
                              const int32_t* ids, int64_t gamma_off, int64_t beta_off, float* bn_stats, int64_t ld_stats,
                              int64_t mean_off, int64_t var_off, float* save_mean, float* save_invstd, float eps,
                              float momentum, int train, int act, float slope, cgl_stream_t stream) {
  if (G == 0) return CGL_OK;
  CGL_REQUIRE(G > 0 && G <= 65535 && rows > 0 && F > 0, "bad shape G=%d rows=%d F=%d", G, rows, F);
  CGL_REQUIRE(u && h && params && (train || bn_stats), "NULL tensor pointer");
  cudaStream_t st = (cudaStream_t)stream;
  BnFwdParams b = {};
  b.rows = rows; b.F = F;
  b.u = u; b.u_gstride = (long long)rows * F;
  b.h = h; b.h_gstride = (long long)rows * F;
  b.params = params; b.ldp = ldp; b.ids = ids; b.gamma_off = gamma_off; b.beta_off = beta_off;
  b.stats = bn_stats; b.ld_stats = ld_stats; b.mean_off = mean_off; b.var_off = var_off;
  b.save_mean = save_mean; b.save_invstd = save_invstd;
  b.eps = eps; b.momentum = momentum; b.train = train; b.act = act; b.slope = slope;
  dim3 grid((F + 127) / 128, G);
  ProfScope prof(CGL_PROF_BN_FWD, 8.0 * G * rows * (double)F, 0.0, st);
  if (bn_smem_ok(rows, F, b.u, b.u_gstride, nullptr, 0, 1)) {
    static unsigned long long attr = 0;
    if (first_use_on_device(attr))
      CGL_CHECK_CUDA(cudaFuncSetAttribute(bn_fwd_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    bn_fwd_smem_kernel<<<grid, 128, (size_t)rows * 512, st>>>(b);
  } else {
    bn_fwd_kernel<<<grid, 128, 0, st>>>(b);
  }
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

// dz: gradient wrt the BatchNorm OUTPUT, overwritten with the gradient wrt its input; gamma / beta take their Adam step
// (step[row] must already count this update).
extern "C" int cgl_bn_backward(int G, int rows, int F, float* dz, const float* u, const float* save_mean,
                               const float* save_invstd, float* params, float* adam_m, float* adam_v, int64_t ldp,
                               const int32_t* ids, int64_t gamma_off, int64_t beta_off, const int32_t* step, float lr,
                               float beta1, float beta2, float eps, cgl_stream_t stream) {
  if (G == 0) return CGL_OK;
  CGL_REQUIRE(G > 0 && G <= 65535 && rows > 0 && F > 0, "bad shape G=%d rows=%d F=%d", G, rows, F);
  CGL_REQUIRE(dz && u && save_mean && save_invstd && params && adam_m && adam_v && step, "NULL tensor pointer");
  cudaStream_t st = (cudaStream_t)stream;
  BnBwdParams b = {};
  b.rows = rows; b.F = F;
  b.dz = dz; b.dz_gstride = (long long)rows * F;
  b.u = u; b.u_gstride = (long long)rows * F;
  b.save_mean = save_mean; b.save_invstd = save_invstd;
  b.params = params; b.adam_m = adam_m; b.adam_v = adam_v; b.ldp = ldp; b.ids = ids;
  b.gamma_off = gamma_off; b.beta_off = beta_off;
  b.step = step; b.lr = lr; b.b1 = beta1; b.b2 = beta2; b.eps = eps; b.scal = nullptr;
  dim3 grid((F + 127) / 128, G);
  ProfScope prof(CGL_PROF_BN_BWD, 12.0 * G * rows * (double)F, 0.0, st);
  if (bn_smem_ok(rows, F, b.dz, b.dz_gstride, b.u, b.u_gstride, 2)) {
    static unsigned long long attr = 0;
    if (first_use_on_device(attr))
      CGL_CHECK_CUDA(cudaFuncSetAttribute(bn_bwd_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    bn_bwd_smem_kernel<<<grid, 128, (size_t)rows * 1024, st>>>(b);
  } else {
    bn_bwd_kernel<<<grid, 128, 0, st>>>(b);
  }
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

// dz = dy * act'(y), y the saved activation OUTPUT
extern "C" int cgl_act_backward(int64_t n, const float* dy, const float* y, float* dz, int act, float slope,
                                cgl_stream_t stream) {
  if (n == 0) return CGL_OK;
  CGL_REQUIRE(n > 0 && dy && y && dz, "bad arguments");
  CGL_REQUIRE(aligned16(dy) && aligned16(y) && aligned16(dz), "cgl_act_backward needs 16-byte aligned tensors");
  const long long nthreads = (n + 3) / 4;
  act_bwd_kernel<<<(unsigned)((nthreads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, dy, y, dz, act, slope);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}

// BatchNorm backward over TWO row segments that were normalised separately (two forward calls through the same layer:
// net_d(real) and net_d(fake) each use their own batch statistics, but their gradients meet in ONE optimizer step):
// dz [G][rows0 + rows1][F], segment 1 follows segment 0; d gamma / d beta are summed over both, Adam is applied once.
namespace cgl {
__global__ void __launch_bounds__(128) bn_bwd_seg_kernel(int rows0, int rows1, int F, float* dz, const float* __restrict__ u,
                                                        const float* __restrict__ mean0, const float* __restrict__ invstd0,
                                                        const float* __restrict__ mean1, const float* __restrict__ invstd1,
                                                        float* params, float* adam_m, float* adam_v, long long ldp, const int* ids,
                                                        long long gamma_off, long long beta_off, const int* step, float lr, float b1,
                                                        float b2, float eps) {
  const int g = blockIdx.y;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int rowid = ids ? ids[g] : g;
  const long long gs = (long long)(rows0 + rows1) * F;
  float* d = dz + (long long)g * gs + f;
  const float* x = u + (long long)g * gs + f;
  const long long go = (long long)rowid * ldp + gamma_off + f, bo = (long long)rowid * ldp + beta_off + f;
  const float gamma = params[go];
  float dgamma = 0.f, dbeta = 0.f;
  for (int seg = 0; seg < 2; ++seg) {
    const int n = seg ? rows1 : rows0;
    if (n == 0) continue;
    const float mean = (seg ? mean1 : mean0)[(long long)g * F + f];
    const float invstd = (seg ? invstd1 : invstd0)[(long long)g * F + f];
    float* ds = d + (long long)(seg ? rows0 : 0) * F;
    const float* xs = x + (long long)(seg ? rows0 : 0) * F;
    float sb0 = 0.f, sb1 = 0.f, sg0 = 0.f, sg1 = 0.f;
    int r = 0;
    for (; r + 1 < n; r += 2) {
      const float d0 = ds[(long long)r * F], d1 = ds[(long long)(r + 1) * F];
      sb0 += d0; sb1 += d1;
      sg0 = fmaf(d0, xs[(long long)r * F] - mean, sg0);
      sg1 = fmaf(d1, xs[(long long)(r + 1) * F] - mean, sg1);
    }
    for (; r < n; ++r) {
      const float d0 = ds[(long long)r * F];
      sb0 += d0;
      sg0 = fmaf(d0, xs[(long long)r * F] - mean, sg0);
    }
    const float sb = sb0 + sb1, dotp = sg0 + sg1;
    dbeta += sb;
    dgamma += dotp * invstd;
    const float k = dotp * invstd * invstd / (float)n, mb = sb / (float)n, a = invstd * gamma;
    for (r = 0; r < n; ++r) ds[(long long)r * F] = (ds[(long long)r * F] - mb - (xs[(long long)r * F] - mean) * k) * a;
  }
  const AdamScalars s = make_adam_scalars(step[rowid], lr, b1, b2, eps);
  {
    float w = gamma, mm = adam_m[go], vv = adam_v[go];
    adam_update(w, mm, vv, dgamma, s);
    params[go] = w; adam_m[go] = mm; adam_v[go] = vv;
  }
  {
    float w = params[bo], mm = adam_m[bo], vv = adam_v[bo];
    adam_update(w, mm, vv, dbeta, s);
    params[bo] = w; adam_m[bo] = mm; adam_v[bo] = vv;
  }
}
}  // namespace cgl

extern "C" int cgl_bn_backward_seg(int G, int rows0, int rows1, int F, float* dz, const float* u, const float* mean0,
                                   const float* invstd0, const float* mean1, const float* invstd1, float* params,
                                   float* adam_m, float* adam_v, int64_t ldp, const int32_t* ids, int64_t gamma_off,
                                   int64_t beta_off, const int32_t* step, float lr, float beta1, float beta2, float eps,
                                   cgl_stream_t stream) {
  if (G == 0) return CGL_OK;
  CGL_REQUIRE(G > 0 && G <= 65535 && rows0 >= 0 && rows1 >= 0 && rows0 + rows1 > 0 && F > 0, "bad shape");
  CGL_REQUIRE(dz && u && mean0 && invstd0 && (rows1 == 0 || (mean1 && invstd1)) && params && adam_m && adam_v && step,
              "NULL tensor pointer");
  dim3 grid((F + 127) / 128, G);
  ProfScope prof(CGL_PROF_BN_BWD, 12.0 * G * (double)(rows0 + rows1) * F, 0.0, (cudaStream_t)stream);
  cgl::bn_bwd_seg_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(rows0, rows1, F, dz, u, mean0, invstd0, mean1, invstd1, params,
                                                                adam_m, adam_v, ldp, ids, gamma_off, beta_off, step, lr, beta1,
                                                                beta2, eps);
  CGL_CHECK_LAUNCH();
  return CGL_OK;
}
