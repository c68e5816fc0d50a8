// fl.cu -- cgl_fl_step: one local minibatch of an FL-style client (the client owns a generator AND a discriminator)
// as ONE call of the C ABI (row a5 of SURVEY.md section 8; the export the section 8b table names).
// Reference: Worker.train, FLGAN/MNIST/flgan.py:251-269 == FLGAN/2DMG/flgan.py:239-256 == fegan.py:284-303:
//     Xd = net_g(z);  D step on (real, Xd);  Xg = net_g(z');  g_loss = BCE(net_d(Xg), 1);  g_loss.backward();  opti_g.step()
// The generator gradients that D_loss.backward() leaves behind are zeroed at opti_g.zero_grad() (flgan.py:264), so the
// first generator pass is a plain forward (it still updates the BatchNorm running statistics: train mode).
// It is the composition of the path's own entry points on one stream -- the same kernels, in the same order, as the host
// loop used to issue one by one -- with the intermediate batches (Xd, Xg, dLoss/dXg) kept in the caller's workspace.
#include "common.cuh"

extern "C" {
size_t cgl_d_step_workspace_bytes(const cgl_mlp_desc* arch, int G, int B);
size_t cgl_g_loss_workspace_bytes(const cgl_mlp_desc* arch, int G, int B);
size_t cgl_mlp_workspace_bytes(const cgl_mlp_desc* arch, int G, int rows);
}

namespace {
inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }
struct FlWs {
  size_t x_off, dx_off, g_off, d_off, g_bytes, d_bytes, total;
};
FlWs fl_carve(const cgl_mlp_desc* ag, const cgl_mlp_desc* ad, int G, int B) {
  FlWs w;
  const size_t batch = up256((size_t)G * B * ad->dims[0] * sizeof(float));
  w.x_off = 0;                 // Xd, then Xg (the D step has consumed Xd before the second generator pass writes Xg)
  w.dx_off = batch;            // dLoss/dXg
  w.g_off = 2 * batch;
  w.g_bytes = up256(cgl_mlp_workspace_bytes(ag, G, B));
  w.d_off = w.g_off + w.g_bytes;
  const size_t a = cgl_d_step_workspace_bytes(ad, G, B), b = cgl_g_loss_workspace_bytes(ad, G, B);
  w.d_bytes = up256(a > b ? a : b);
  w.total = w.d_off + w.d_bytes;
  return w;
}
}  // namespace

extern "C" size_t cgl_fl_step_workspace_bytes(const cgl_mlp_desc* arch_g, const cgl_mlp_desc* arch_d, int G, int B) {
  if (!arch_g || !arch_d || G <= 0 || B <= 0) return 0;
  return fl_carve(arch_g, arch_d, G, B).total;
}

extern "C" int cgl_fl_step(const cgl_mlp_desc* arch_g, const cgl_mlp_desc* arch_d, int G, float* g_params, float* g_adam_m,
                           float* g_adam_v, int64_t ld_g, int32_t* g_step, float* g_bn_stats, int64_t ld_stats,
                           float* d_params, float* d_adam_m, float* d_adam_v, int64_t ld_d, int32_t* d_step,
                           const int32_t* ids, const float* z_d, const float* z_g, const float* real,
                           const int32_t* n_real, int B, const cgl_train_cfg* cfg_d, const cgl_train_cfg* cfg_g,
                           float* out_dloss, float* out_gloss, void* workspace, size_t workspace_bytes,
                           cgl_stream_t stream) {
  CGL_REQUIRE(arch_g && arch_d && cfg_d && cfg_g, "arch / cfg is NULL");
  if (G == 0) return CGL_OK;
  CGL_REQUIRE(G > 0 && B > 0, "bad shape G=%d B=%d", G, B);
  CGL_REQUIRE(z_d && z_g && real && out_dloss && out_gloss && workspace, "NULL tensor pointer");
  CGL_REQUIRE(arch_g->dims[arch_g->n_layers] == arch_d->dims[0],
              "the generator's output width (%d) is not the discriminator's input width (%d)",
              arch_g->dims[arch_g->n_layers], arch_d->dims[0]);
  const FlWs w = fl_carve(arch_g, arch_d, G, B);
  if (workspace_bytes < w.total) {
    cgl::set_error("workspace too small: %zu < %zu", workspace_bytes, w.total);
    return CGL_EWORKSPACE;
  }
  char* base = (char*)workspace;
  float* X = (float*)(base + w.x_off);
  float* dX = (float*)(base + w.dx_off);
  void* gws = base + w.g_off;
  void* dws = base + w.d_off;
  const int zdim = arch_g->dims[0];
  int rc;
  // Xd = net_g(z)                                                         flgan.py:251-252
  rc = cgl_mlp_forward(arch_g, G, g_params, ld_g, ids, g_bn_stats, ld_stats, 1, z_d, (int64_t)B * zdim, nullptr, B, X, gws,
                       w.g_bytes, stream);
  if (rc) return rc;
  // D_loss = BCE(net_d(real), 1) + BCE(net_d(Xd), 0); backward; opti_d.step()        flgan.py:255-261
  rc = cgl_d_step(arch_d, G, d_params, d_adam_m, d_adam_v, ld_d, d_step, ids, real, n_real, X, nullptr, B, cfg_d, out_dloss,
                  dws, w.d_bytes, stream);
  if (rc) return rc;
  // Xg = net_g(z'); g_loss = BCE(net_d(Xg), 1)                             flgan.py:263-267
  rc = cgl_mlp_forward(arch_g, G, g_params, ld_g, ids, g_bn_stats, ld_stats, 1, z_g, (int64_t)B * zdim, nullptr, B, X, gws,
                       w.g_bytes, stream);
  if (rc) return rc;
  rc = cgl_g_loss(arch_d, G, d_params, ld_d, ids, X, nullptr, B, cfg_d->loss_kind, out_gloss, dX, dws, w.d_bytes, stream);
  if (rc) return rc;
  // g_loss.backward(); opti_g.step()                                       flgan.py:268-269
  return cgl_mlp_backward(arch_g, G, g_params, g_adam_m, g_adam_v, ld_g, g_step, ids, cfg_g, z_g, (int64_t)B * zdim, nullptr,
                          B, X, dX, nullptr, gws, w.g_bytes, stream);
}
