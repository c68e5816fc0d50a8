// arch.cu -- host-only part of the ABI: error text, architecture tables, packed-row layout.
// Shapes follow the reference model classes; nothing here touches the GPU
// except cgl_device_ok().
#include "common.cuh"
#include <string.h>
#include <vector>

namespace cgl {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static long long g_launches = 0;
void count_launch(int n) { g_launches += n; }

struct ProfRec { int tag; cudaEvent_t a, b; double bytes, flops; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_pool;   // recycled events
static cudaEvent_t prof_event() {
  cudaEvent_t e;
  if (!g_prof_pool.empty()) { e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEventCreate(&e);
  return e;
}
void prof_begin(int tag, double bytes, double flops, cudaStream_t st) {
  if (!g_prof_on) return;
  ProfRec r = {tag, prof_event(), prof_event(), bytes, flops};
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t st) {
  if (!g_prof_on || g_prof.empty()) return;
  cudaEventRecord(g_prof.back().b, st);
}
}  // namespace cgl

using namespace cgl;

extern "C" long long cgl_launch_count(void) { return g_launches; }

static const char* kProfNames[CGL_PROF_NUM_TAGS] = {
    "linear_fwd[tcgen05]", "linear_bwd_data[tcgen05]", "linear_wgrad+adam[tcgen05]", "linear_wgrad[tcgen05]",
    "linear_fwd[ffma]", "linear_bwd_data[ffma]", "linear_wgrad+adam[ffma]", "linear_wgrad[ffma]",
    "head_loss", "batchnorm_fwd", "batchnorm_bwd", "mix/aggregate", "elementwise", "client_step_fused[ffma]"};

extern "C" int cgl_profile_enable(int on) {
  for (auto& r : g_prof) { g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b); }
  g_prof.clear();
  g_prof_on = on != 0;
  return CGL_OK;
}
extern "C" const char* cgl_profile_tag_name(int tag) {
  return (tag >= 0 && tag < CGL_PROF_NUM_TAGS) ? kProfNames[tag] : "";
}
extern "C" int cgl_profile_summary(int tag, double* out_ms, double* out_bytes, double* out_flops, long long* out_n) {
  CGL_REQUIRE(tag >= 0 && tag < CGL_PROF_NUM_TAGS, "bad profile tag %d", tag);
  double ms = 0, by = 0, fl = 0;
  long long n = 0;
  for (auto& r : g_prof) {
    if (r.tag != tag) continue;
    CGL_CHECK_CUDA(cudaEventSynchronize(r.b));
    float t = 0.f;
    CGL_CHECK_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms += t; by += r.bytes; fl += r.flops; ++n;
  }
  if (out_ms) *out_ms = ms;
  if (out_bytes) *out_bytes = by;
  if (out_flops) *out_flops = fl;
  if (out_n) *out_n = n;
  return CGL_OK;
}

extern "C" const char* cgl_version(void) { return "cgl_b200 0.1.0 (sm_100a)"; }
extern "C" const char* cgl_last_error(void) { return g_err; }

extern "C" int cgl_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return 0;
  }
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

static void fill(cgl_mlp_desc* d, int n, const int* dims, const int* act, const int* bn) {
  memset(d, 0, sizeof(*d));
  d->n_layers = n;
  for (int i = 0; i <= n; ++i) d->dims[i] = dims[i];
  for (int i = 0; i < n; ++i) {
    d->act[i] = act[i];
    d->bn[i] = bn ? bn[i] : 0;
  }
  d->bn_eps = 0.8f;       // nn.BatchNorm1d(out_feat, 0.8): model/mnist_model.py:13
  d->bn_momentum = 0.1f;  // torch default
  d->lrelu_slope = 0.2f;  // nn.LeakyReLU(0.2)
}

extern "C" int cgl_arch_describe(int arch_id, cgl_mlp_desc* out) {
  CGL_REQUIRE(out != nullptr, "out_desc is NULL");
  const int L = CGL_ACT_LRELU, T = CGL_ACT_TANH, S = CGL_ACT_SIGMOID, N = CGL_ACT_NONE;
  switch (arch_id) {
    case CGL_ARCH_D_2D: {  // CGLGAN/2DMG/model.py:58-66
      int dims[] = {2, 128, 256, 1}; int act[] = {L, L, S};
      fill(out, 3, dims, act, nullptr); break;
    }
    case CGL_ARCH_D_MNIST1: {  // CGLGAN/MNIST/mnist_model.py:74-81
      int dims[] = {784, 512, 256, 1}; int act[] = {L, L, S};
      fill(out, 3, dims, act, nullptr); break;
    }
    case CGL_ARCH_D_MNIST2: {  // model/mnist_model.py:76-83
      int dims[] = {784, 512, 256, 2}; int act[] = {L, L, N};
      fill(out, 3, dims, act, nullptr); break;
    }
    case CGL_ARCH_D_MNIST_LS: {
      int dims[] = {784, 512, 256, 1}; int act[] = {L, L, N};
      fill(out, 3, dims, act, nullptr); break;
    }
    case CGL_ARCH_G_2D_MD: {  // MDGAN/2DMG/model.py:8-15
      int dims[] = {100, 256, 128, 2}; int act[] = {L, L, T};
      fill(out, 3, dims, act, nullptr); break;
    }
    case CGL_ARCH_G_MNIST: {  // model/mnist_model.py:17-24
      int dims[] = {100, 128, 256, 512, 1024, 784}; int act[] = {L, L, L, L, T}; int bn[] = {0, 1, 1, 1, 0};
      fill(out, 5, dims, act, bn); break;
    }
    case CGL_ARCH_G_2D_TRUNK: {  // CGLGAN/2DMG/model.py:30-33
      int dims[] = {100, 32}; int act[] = {L};
      fill(out, 1, dims, act, nullptr); break;
    }
    case CGL_ARCH_G_2D_HEAD: {  // CGLGAN/2DMG/model.py:36-41
      int dims[] = {32, 2}; int act[] = {T};
      fill(out, 1, dims, act, nullptr); break;
    }
    case CGL_ARCH_G_MNIST_TRUNK: {  // model/mnist_model.py:45-49
      int dims[] = {100, 128, 256, 512}; int act[] = {L, L, L}; int bn[] = {0, 1, 1};
      fill(out, 3, dims, act, bn); break;
    }
    case CGL_ARCH_G_MNIST_HEAD: {  // model/mnist_model.py:52-57
      int dims[] = {512, 1024, 784}; int act[] = {L, T}; int bn[] = {1, 0};
      fill(out, 2, dims, act, bn); break;
    }
    default:
      set_error("unknown arch id %d", arch_id);
      return CGL_EINVAL;
  }
  return CGL_OK;
}

// parameters() order of nn.Sequential(Linear, [BatchNorm1d], act, ...): weight, bias, [bn.weight, bn.bias]
extern "C" int cgl_mlp_layout_of(const cgl_mlp_desc* d, cgl_mlp_layout* out) {
  CGL_REQUIRE(d && out, "NULL argument");
  CGL_REQUIRE(d->n_layers >= 1 && d->n_layers <= CGL_MAX_LAYERS, "n_layers=%d out of range", d->n_layers);
  memset(out, 0, sizeof(*out));
  int64_t off = 0, soff = 0;
  for (int i = 0; i < CGL_MAX_LAYERS; ++i) {
    out->w_off[i] = out->b_off[i] = out->bn_w_off[i] = out->bn_b_off[i] = -1;
    out->bn_mean_off[i] = out->bn_var_off[i] = -1;
  }
  for (int i = 0; i < d->n_layers; ++i) {
    CGL_REQUIRE(d->dims[i] > 0 && d->dims[i + 1] > 0, "bad width at layer %d", i);
    out->w_off[i] = off; off += (int64_t)d->dims[i] * d->dims[i + 1];
    out->b_off[i] = off; off += d->dims[i + 1];
    if (d->bn[i]) {
      out->bn_w_off[i] = off; off += d->dims[i + 1];
      out->bn_b_off[i] = off; off += d->dims[i + 1];
      out->bn_mean_off[i] = soff; soff += d->dims[i + 1];
      out->bn_var_off[i] = soff; soff += d->dims[i + 1];
    }
  }
  out->n_params = off;
  out->n_bn_stats = soff;
  return CGL_OK;
}
