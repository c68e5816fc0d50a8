/*
 * cgl_b200.h -- C ABI of the B200 (sm_100a) engine for the CGL-GAN simulated-client hot path.
 *
 * The reference (NetworkCommunication/CGL-GAN) is pure Python/PyTorch and has no FFI of its own;
 * its "operator interface" for this path is the body of Worker.train / Server.train / Cloud.run.
 * Each entry point below names the reference code it replaces (file:line under /root/reference).
 *
 * Conventions
 *   - every function returns 0 on success and a negative CGL_E* code on failure; it never throws
 *     and never allocates device memory. cgl_last_error() gives a human-readable reason.
 *   - all tensor pointers are CALLER-OWNED DEVICE pointers (fp32 unless stated, indices int32).
 *   - every compute call is asynchronous on the cudaStream_t passed as `stream` (cgl_stream_t).
 *   - "packed row": one client's parameters as a flat fp32 vector in torch `parameters()` order
 *     (= fedlab SerializationTool.serialize_model order, capgan.py:170), rows `ldp` floats apart.
 */
#ifndef CGL_B200_H
#define CGL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cgl_stream_t; /* a cudaStream_t */
typedef void* cgl_comm_t;   /* opaque communicator (wraps an ncclComm_t) */

/* ---- error codes ---- */
#define CGL_OK 0
#define CGL_EINVAL (-1)     /* bad argument (NULL pointer, unsupported arch/loss pairing, ...) */
#define CGL_EWORKSPACE (-2) /* workspace too small */
#define CGL_ECUDA (-3)      /* a CUDA runtime call / kernel launch failed */
#define CGL_ENCCL (-4)      /* NCCL unavailable or an NCCL call failed */
#define CGL_EUNSUPPORTED (-5)

/* ---- activations / losses ---- */
#define CGL_ACT_NONE 0
#define CGL_ACT_LRELU 1 /* LeakyReLU(0.2): CGLGAN/2DMG/model.py:60 */
#define CGL_ACT_TANH 2
#define CGL_ACT_SIGMOID 3

#define CGL_LOSS_BCE 0 /* nn.BCELoss on a sigmoid output: CGLGAN/2DMG/main.py:336,361-364 */
#define CGL_LOSS_CE 1  /* nn.CrossEntropyLoss on 2 logits: capgan.py:324-339 */
#define CGL_LOSS_MSE 2 /* LSGAN least-squares (no call site in the reference; parity unpinned) */

/* ---- architectures known to cgl_arch_describe ---- */
#define CGL_ARCH_D_2D 0      /* 2->128->256->1 sigmoid : CGLGAN/2DMG/model.py:54-71 */
#define CGL_ARCH_D_MNIST1 1  /* 784->512->256->1 sigmoid: CGLGAN/MNIST/mnist_model.py:69-86 */
#define CGL_ARCH_D_MNIST2 2  /* 784->512->256->2 logits : model/mnist_model.py:71-88 */
#define CGL_ARCH_D_MNIST_LS 3 /* 784->512->256->1 linear (LSGAN/MSE variant of D_MNIST1) */
#define CGL_ARCH_G_2D_MD 4   /* 100->256->128->2 tanh   : MDGAN/2DMG/model.py:4-21 */
#define CGL_ARCH_G_MNIST 5   /* 100->128->256bn->512bn->1024bn->784 tanh: model/mnist_model.py:5-29 */
#define CGL_ARCH_G_2D_TRUNK 6  /* 100->32 lrelu           : CGLGAN/2DMG/model.py:30-33 */
#define CGL_ARCH_G_2D_HEAD 7   /* 32->2 tanh              : CGLGAN/2DMG/model.py:36-41 */
#define CGL_ARCH_G_MNIST_TRUNK 8 /* 100->128->256bn->512bn : model/mnist_model.py:45-49 */
#define CGL_ARCH_G_MNIST_HEAD 9  /* 512->1024bn->784 tanh  : model/mnist_model.py:52-57 */
#define CGL_NUM_ARCH 10

#define CGL_MAX_LAYERS 8

/* A stack of Linear[+BatchNorm1d]+activation layers. */
typedef struct cgl_mlp_desc {
  int32_t n_layers;                 /* number of Linear layers */
  int32_t dims[CGL_MAX_LAYERS + 1]; /* dims[0] = input width, dims[i+1] = out features of layer i */
  int32_t act[CGL_MAX_LAYERS];      /* CGL_ACT_* applied after layer i (after BN if bn[i]) */
  int32_t bn[CGL_MAX_LAYERS];       /* 1: BatchNorm1d(out, eps=bn_eps) between Linear i and act */
  float bn_eps;                     /* 0.8 in the reference (2nd positional arg of BatchNorm1d) */
  float bn_momentum;                /* 0.1 */
  float lrelu_slope;                /* 0.2 */
} cgl_mlp_desc;

/* Offsets (in floats) of each parameters() entry inside a packed row. */
typedef struct cgl_mlp_layout {
  int64_t n_params;                   /* P: total floats in a packed row */
  int64_t w_off[CGL_MAX_LAYERS];      /* Linear.weight [out,in] row-major */
  int64_t b_off[CGL_MAX_LAYERS];      /* Linear.bias   [out] */
  int64_t bn_w_off[CGL_MAX_LAYERS];   /* BatchNorm.weight [out], -1 if none */
  int64_t bn_b_off[CGL_MAX_LAYERS];   /* BatchNorm.bias   [out], -1 if none */
  int64_t n_bn_stats;                 /* floats of running_mean+running_var (kept in a second row) */
  int64_t bn_mean_off[CGL_MAX_LAYERS]; /* offset inside the stats row, -1 if none */
  int64_t bn_var_off[CGL_MAX_LAYERS];
} cgl_mlp_layout;

/* Optimiser / loss knobs of one discriminator step. */
typedef struct cgl_train_cfg {
  int32_t loss_kind;  /* CGL_LOSS_* */
  float d_loss_scale; /* 0.5 in capgan.py:339 / mixed-gan.py:382, 1.0 elsewhere */
  float lr;           /* lr_d = 2e-4: CGLGAN/2DMG/main.py:284 */
  float beta1, beta2; /* b1=0.5, b2=0.999: CGLGAN/2DMG/main.py:55-56 */
  float eps;          /* torch.optim.Adam default 1e-8 */
} cgl_train_cfg;

/* ---- introspection ---- */
const char* cgl_version(void);
const char* cgl_last_error(void);
/* 1 if a CUDA device with compute capability 10.x is usable, 0 otherwise (never fails). */
int cgl_device_ok(void);
/* Number of CUDA kernels this library has launched in this process (bench.py's gpu_launches). */
long long cgl_launch_count(void);

/* Replaces: the model constructors (a6) as far as shapes go. */
int cgl_arch_describe(int arch_id, cgl_mlp_desc* out_desc);
/* Replaces: SerializationTool.serialize_model ordering (capgan.py:170, fegan.py:133-134). */
int cgl_mlp_layout_of(const cgl_mlp_desc* desc, cgl_mlp_layout* out_layout);

/* ---- K1/K2: the per-client discriminator step -------------------------------------------
 * Replaces Worker.train's D loop: CGLGAN/2DMG/main.py:349-366 (BCE), capgan.py:324-341 (CE x0.5),
 * MDGAN/MNIST/mdgan.py:270-288, ACGAN/MNIST/acgan.py:273-292, and the Adam step (a7).
 * For each of the G groups (clients) g:
 *     row   = client_ids ? client_ids[g] : g                 (row of params/adam_m/adam_v/step)
 *     L     = loss(D(real[g][0:n_real[g]]), valid) + loss(D(fake[fake_idx ? fake_idx[g] : g]), fake)
 *     out_dloss[g] = d_loss_scale * L ; backward ; step[row] += 1 ; Adam(row)
 * real is [G, B, d] (rows >= n_real[g] are ignored), fake is [*, B, d]; n_real NULL means all B.
 * workspace must hold cgl_d_step_workspace_bytes(arch, G, B) bytes.                       */
size_t cgl_d_step_workspace_bytes(const cgl_mlp_desc* arch, int G, int B);
int cgl_d_step(const cgl_mlp_desc* arch, int G, float* params, float* adam_m, float* adam_v,
               int64_t ldp, int32_t* step, const int32_t* client_ids, const float* real,
               const int32_t* n_real, const float* fake, const int32_t* fake_idx, int B,
               const cgl_train_cfg* cfg, float* out_dloss, void* workspace, size_t workspace_bytes,
               cgl_stream_t stream);

/* ---- K1/K2: generator-loss evaluation through each client's (updated) D -------------------
 * Replaces Worker.train's tail: CGLGAN/2DMG/main.py:368-373, capgan.py:343-347 -- and the part of
 * Server.train's backward that runs through every client's D (CGLGAN/2DMG/main.py:257,268):
 *     out_loss[g] = loss(D_row(xg[xg_idx ? xg_idx[g] : g]), valid)
 *     out_dxg[g]  = d out_loss[g] / d xg        ([G, B, d]; may be NULL)
 * This is the (client_idx, F_grad, F_pred) hand-off sketched at CGLGAN/MNIST/main.py:220-235.  */
size_t cgl_g_loss_workspace_bytes(const cgl_mlp_desc* arch, int G, int B);
int cgl_g_loss(const cgl_mlp_desc* arch, int G, const float* params, int64_t ldp,
               const int32_t* client_ids, const float* xg, const int32_t* xg_idx, int B,
               int loss_kind, float* out_loss, float* out_dxg, void* workspace,
               size_t workspace_bytes, cgl_stream_t stream);

/* ---- K1: the whole client step of a round in ONE call ---------------------------------------
 * Replaces one Worker.train call with epoch == 1 and one server: CGLGAN/2DMG/main.py:344-375 (the D loop :357-366,
 * then the tail :368-373 through the UPDATED D), MDGAN/2DMG/mdgan.py:252-278, ACGAN/2DMG/acgan.py:231-257.
 * = cgl_d_step followed by cgl_g_loss on the same rows. For the small 2DMG discriminator
 * (CGLGAN/2DMG/model.py:54-71: d <= 4 -> 128 -> 256 -> 1, BCE or MSE) this is ONE kernel launch with one CTA per
 * client: the client's weights are read from HBM once, stay in shared memory through forward, loss, backward, the
 * fused Adam update and the generator-loss pass (csrc/client_fused.cuh); cgl_d_step and cgl_g_loss take the same
 * kernel for these networks when called on their own. Other discriminators run the layered kernels.
 * workspace: cgl_d_step_workspace_bytes(arch, G, B) bytes.                                          */
int cgl_client_step(const cgl_mlp_desc* arch, int G, float* params, float* adam_m, float* adam_v, int64_t ldp,
                    int32_t* step, const int32_t* client_ids, const float* real, const int32_t* n_real,
                    const float* fake, const int32_t* fake_idx, const float* xg, const int32_t* xg_idx, int B,
                    const cgl_train_cfg* cfg, float* out_dloss, float* out_gloss, float* out_dxg,
                    void* workspace, size_t workspace_bytes, cgl_stream_t stream);

/* out[s] = sum_{j in [srv_ptr[s], srv_ptr[s+1])} weights[clients[j]] * dxg[clients[j]]
 * (weights NULL = 1, clients NULL = identity). Replaces the accumulation of every client's
 * dLoss/dXg into a shared Xg during F_max.backward(): capgan.py:258, MDGAN/MNIST/mdgan.py:203-204. */
int cgl_dxg_reduce(int S, const int32_t* srv_ptr, const int32_t* clients, const float* weights,
                   const float* dxg, int64_t n /* B*d floats per client */, float* out,
                   cgl_stream_t stream);

/* ---- the generator side of a round (a4, a5) ---------------------------------------------------------
 * A stack of Linear [+ BatchNorm1d(eps 0.8, batch statistics)] + LeakyReLU/Tanh/Sigmoid layers over G
 * independent groups (edge servers, generator heads, FL clients), parameters in packed rows.
 *   row  = ids ? ids[g] : g                      (row of params / adam_* / step / bn_stats)
 *   x[g] = x + (x_idx ? x_idx[g] : g) * x_gstride,  [rows, dims[0]]     (heads read their server's trunk output)
 * cgl_mlp_forward: y[g] = net_row(x[g]), [G, rows, dims[L]]. train != 0: BatchNorm uses the batch statistics and
 *   updates running_mean / running_var in bn_stats (both generator passes of Server.train do,
 *   CGLGAN/2DMG/main.py:229-234); train == 0: running statistics (G.eval() snapshots, :217-223).
 *   The activations needed by the backward stay in `workspace` until the next forward on it.
 * cgl_mlp_backward: given dy = dLoss/dy of the LAST forward on `workspace`, back-propagates through every layer
 *   (model/mnist_model.py:17-24), takes one Adam step on every parameter of the row (opti_g.step(),
 *   CGLGAN/2DMG/main.py:276; step[row] += 1) and, if dx != NULL, writes dLoss/dx [G, rows, dims[0]]
 *   (the heads' gradient into the shared trunk, mixed-gan.py:263-281).
 * Replaces: Xd = net_g(z) / Xg = net_g(z) (CGLGAN/2DMG/main.py:229-234, mixed-gan.py:241-252), the generator part
 * of F_max.backward() and opti.step() (:266-276); FL: FLGAN/MNIST/flgan.py:251-252,263-269.                  */
size_t cgl_mlp_workspace_bytes(const cgl_mlp_desc* arch, int G, int rows);
int cgl_mlp_forward(const cgl_mlp_desc* arch, int G, const float* params, int64_t ldp, const int32_t* ids,
                    float* bn_stats, int64_t ld_stats, int train, const float* x, int64_t x_gstride,
                    const int32_t* x_idx, int rows, float* y, void* workspace, size_t workspace_bytes,
                    cgl_stream_t stream);
int cgl_mlp_backward(const cgl_mlp_desc* arch, int G, float* params, float* adam_m, float* adam_v, int64_t ldp,
                     int32_t* step, const int32_t* ids, const cgl_train_cfg* cfg /* lr, beta1, beta2, eps */,
                     const float* x, int64_t x_gstride, const int32_t* x_idx, int rows, const float* y,
                     const float* dy, float* dx, void* workspace, size_t workspace_bytes, cgl_stream_t stream);

/* ---- one FL-style local minibatch in one call (a5) -------------------------------------------------------
 * Replaces the body of FL Worker.train's minibatch loop, FLGAN/MNIST/flgan.py:251-269, FLGAN/2DMG/flgan.py:239-256,
 * fegan.py:284-303, for G clients that each own a generator (arch_g, BatchNorm statistics in g_bn_stats) and a
 * discriminator (arch_d):   Xd = G(z_d);  D step on (real, Xd) [cgl_d_step];  Xg = G(z_g);  g_loss = loss(D(Xg), valid);
 *                           g_loss.backward();  opti_g.step() [cgl_mlp_backward].
 * ids (NULL: identity) picks the rows of BOTH banks (FeGAN's group of the round). z_d / z_g are [G, B, arch_g->dims[0]],
 * real is [G, B, d] with n_real valid rows. Every intermediate batch stays in `workspace`
 * (cgl_fl_step_workspace_bytes). out_dloss / out_gloss: [G].                                                   */
size_t cgl_fl_step_workspace_bytes(const cgl_mlp_desc* arch_g, const cgl_mlp_desc* arch_d, int G, int B);
int cgl_fl_step(const cgl_mlp_desc* arch_g, const cgl_mlp_desc* arch_d, int G, float* g_params, float* g_adam_m,
                float* g_adam_v, int64_t ld_g, int32_t* g_step, float* g_bn_stats, int64_t ld_stats, float* d_params,
                float* d_adam_m, float* d_adam_v, int64_t ld_d, int32_t* d_step, const int32_t* ids, const float* z_d,
                const float* z_g, const float* real, const int32_t* n_real, int B, const cgl_train_cfg* cfg_d,
                const cgl_train_cfg* cfg_g, float* out_dloss, float* out_gloss, void* workspace, size_t workspace_bytes,
                cgl_stream_t stream);

/* ---- fused Adam over packed rows (server-side G, a7) ------------------------------------
 * torch.optim.Adam(betas=(b1,b2)) semantics, CGLGAN/2DMG/main.py:192. step[r] is incremented.
 * rows: R rows of n floats, `ld` floats apart, for p / g / m / v alike.                      */
int cgl_adam_rows(int R, int64_t n, int64_t ld, float* p, const float* g, float* m, float* v,
                  int32_t* step, float lr, float beta1, float beta2, float eps,
                  cgl_stream_t stream);

/* ---- K3: aggregation over packed parameter rows -------------------------------------------
 * cgl_mix_csr:   dst[r,:] = sum_j vals[j] * src[col[j],:]  for j in [row_ptr[r], row_ptr[r+1])
 *   One op for FedAvg (FLGAN/MNIST/flgan.py:143-163), MD-GAN swap = permutation
 *   (MDGAN/MNIST/mdgan.py:158-164), neighbour / group mean (ACGAN/MNIST/acgan.py:240-263,
 *   CGLGAN/2DMG/main.py:171-179). src and dst must not overlap. Accumulation runs in column order
 *   j, each product and sum rounded to fp32 separately (the dict loop of Cloud.run).
 * cgl_wsum:      out[:] = sum_c w[c] * src[rows ? rows[c] : c, :]   (Cloud.run,
 *   CGLGAN/2DMG/main.py:126-133; fedavg_aggregate, capgan.py:114)
 * cgl_bcast_mix: dst[rows ? rows[r] : r, :] = sigma * dst[..] + (1-sigma) * g[:]   (the `segema`
 *   mix + load_state_dict, CGLGAN/2DMG/main.py:205-208; Worker.run load, flgan.py:222-223)       */
int cgl_mix_csr(int R, int64_t n, const int32_t* row_ptr, const int32_t* col, const float* vals,
                const float* src, int64_t ld_src, float* dst, int64_t ld_dst, cgl_stream_t stream);
int cgl_wsum(int C, int64_t n, const float* w, const int32_t* rows, const float* src, int64_t ld_src,
             float* out, cgl_stream_t stream);
int cgl_bcast_mix(int R, int64_t n, const int32_t* rows, float sigma, const float* g, float* dst,
                  int64_t ld_dst, cgl_stream_t stream);
/* The same sums where the reference DIVIDES instead of multiplying by a pre-computed share (bit-exact for counts that
 * are not powers of two):
 *   sum_first == 0: out[:] = sum_c (src[row(c), :] / divisor)     FL Server.run, `p[key] += paras[key] / len(client_list)`
 *                                                                  (FLGAN/MNIST/flgan.py:151-158, FLGAN/2DMG/flgan.py:149-156)
 *   sum_first != 0: out[:] = (sum_c src[row(c), :]) / divisor     receive_parameter (CGLGAN/2DMG/main.py:171-179)
 * cgl_mix_csr with vals == NULL computes the second form per output row (divisor = the row's column count): the
 * neighbour / group mean of discriminators.                                                                   */
int cgl_wsum_div(int C, int64_t n, float divisor, int sum_first, const int32_t* rows, const float* src,
                 int64_t ld_src, float* out, cgl_stream_t stream);

/* ---- data side path kept on the GPU ----------------------------------
 * cgl_gather_rows: out[i,:] = data[idx[i],:] (idx[i] < 0: a row of zeros = the padding of a short last batch).
 *   The dataset stays resident in HBM; the host reproduces DataLoader(shuffle=True)'s indices
 *   (Worker.__init__ / Worker.train, CGLGAN/2DMG/main.py:299-301,350-355; the unshuffled full pass of
 *   FLGAN/MNIST/flgan.py:250) and ships 8 bytes per sample instead of the sample.
 * cgl_hist2d / cgl_kl_score_2d: plot_2d's quality score (CGLGAN/2DMG/main.py:68-94): np.histogram2d with
 *   16 x 16 bins over [-1,1]^2 of n points (x at xy[i*stride], y at xy[i*stride+1]) and
 *   scipy.stats.entropy(generated, real) over the bins the real set occupies (double, *out_kl on the device). */
int cgl_gather_rows(int64_t n_out, int d, const float* data, int64_t n_rows, const int64_t* idx, float* out,
                    cgl_stream_t stream);
int cgl_hist2d(int64_t n, const float* xy, int64_t stride, uint32_t* hist256, cgl_stream_t stream);
int cgl_kl_score_2d(int64_t n, const float* xy, int64_t stride, const uint32_t* real_hist256,
                    uint32_t* scratch_hist256, double* out_kl, cgl_stream_t stream);

/* ---- cross-GPU aggregation (the only collective on the path) --------------------------------
 * One process per GPU. rank 0 calls cgl_comm_unique_id, ships the 128 bytes to the other ranks
 * (torch.distributed broadcast), every rank calls cgl_comm_init. cgl_mix_allreduce computes the
 * local weighted partial sum (cgl_wsum with pre-normalised weights) and all-reduces it in place
 * over NVLink on `stream`: the Cloud / FL server aggregation when clients are sharded.          */
int cgl_comm_unique_id(uint8_t out_id[128]);
int cgl_comm_init(int nranks, int rank, const uint8_t id[128], cgl_comm_t* out_comm);
int cgl_comm_destroy(cgl_comm_t comm);
int cgl_allreduce_sum(cgl_comm_t comm, float* buf, int64_t n, cgl_stream_t stream);
int cgl_mix_allreduce(cgl_comm_t comm, int C_local, int64_t n, const float* w_local,
                      const int32_t* rows, const float* src, int64_t ld_src, float* out,
                      cgl_stream_t stream);

/* ---- live kernel timing (bench.py's roofline) ---------------------------------------------------------
 * cgl_profile_enable(1) makes every instrumented kernel class record a CUDA-event pair on its launching stream
 * together with its ALGORITHMIC bytes and FLOPs; cgl_profile_summary(tag, ...) synchronises
 * those events and returns the totals since the last cgl_profile_enable call. Off by default.            */
#define CGL_PROF_FWD_TC 0
#define CGL_PROF_BWD_TC 1
#define CGL_PROF_WGRAD_ADAM_TC 2
#define CGL_PROF_WGRAD_TC 3
#define CGL_PROF_FWD_FFMA 4
#define CGL_PROF_BWD_FFMA 5
#define CGL_PROF_WGRAD_ADAM_FFMA 6
#define CGL_PROF_WGRAD_FFMA 7
#define CGL_PROF_HEAD 8
#define CGL_PROF_BN_FWD 9
#define CGL_PROF_BN_BWD 10
#define CGL_PROF_MIX 11
#define CGL_PROF_ELEMENTWISE 12
#define CGL_PROF_CLIENT_FUSED 13
#define CGL_PROF_NUM_TAGS 14
int cgl_profile_enable(int on);
const char* cgl_profile_tag_name(int tag);
int cgl_profile_summary(int tag, double* out_ms, double* out_bytes, double* out_flops, long long* out_launches);

/* ---- kernel selection (tests / profiling) ------------------------------------------------------
 * The Linear products run on one of two sm_100a kernels of this library, chosen by shape and alignment:
 * the tcgen05/TMEM 3xTF32 grouped GEMM (wide, 16-byte aligned layers) or the exact-fp32 FFMA grouped GEMM
 * (narrow or unaligned layers). mode: 0 = automatic (default), 1 = FFMA only, 2 = tcgen05 wherever the
 * operands are addressable by it. Process-wide; not a reference knob.                                */
#define CGL_GEMM_AUTO 0
#define CGL_GEMM_FFMA 1
#define CGL_GEMM_TC 2
int cgl_set_gemm_mode(int mode);
int cgl_get_gemm_mode(void);
/* The fused shared-memory-resident client step (cgl_client_step, csrc/client_fused.cuh) on / off for the networks it
 * covers; off = the layered kernels. Process-wide; CGL_K1=0|1 in the environment presets it. Not a reference knob. */
int cgl_set_fused_client_step(int on);
int cgl_get_fused_client_step(void);
/* Bring-up only: per-CTA clock64() milestones of the tcgen05 GEMM (csrc/tc_gemm.cuh); NULL switches it off. */
int cgl_debug_set_timeline(long long* device_buf);

/* ---- building blocks (exported for tests and for the host-side generator step) -------------
 * Grouped Linear over G independent groups, fp32:
 *   fwd : y[g] = act(x[g] W[g]^T + b[g])        x [rows,in] (ldx), W [out,in], y [rows,out]
 *   bwd : dx[g] = (dy[g] W[g]) * act'(saved[g]) (saved NULL -> no activation derivative)
 *   wgrad: dW[g] = dy[g]^T x[g], db[g] = colsum(dy[g])
 * W/b/dW/db live in packed rows: pointer = base + (ids ? ids[g] : g) * ld + offset.             */
int cgl_linear_fwd(int G, int rows, int in, int out, const float* x, int64_t x_gstride,
                   const float* wbase, int64_t ldp, const int32_t* ids, int64_t w_off, int64_t b_off,
                   int act, float slope, float* y, int64_t y_gstride, cgl_stream_t stream);
int cgl_linear_bwd_data(int G, int rows, int in, int out, const float* dy, int64_t dy_gstride,
                        const float* wbase, int64_t ldp, const int32_t* ids, int64_t w_off,
                        const float* saved, int64_t saved_gstride, int act, float slope, float* dx,
                        int64_t dx_gstride, cgl_stream_t stream);
int cgl_linear_wgrad(int G, int rows, int in, int out, const float* dy, int64_t dy_gstride,
                     const float* x, int64_t x_gstride, float* gbase, int64_t ldg,
                     const int32_t* ids, int64_t w_off, int64_t b_off, cgl_stream_t stream);

/* Weight gradient with the Adam step fused into the epilogue (what cgl_d_step / cgl_mlp_backward run per layer):
 *   W[g] <- Adam(W[g], dy[g]^T x[g]),  b[g] <- Adam(b[g], colsum(dy[g]))   with torch.optim.Adam semantics (a7).
 * step[row(g)] must already hold the step count INCLUDING this update (the bias corrections use it).
 * adam_scratch: NULL, or 32 * G bytes of device memory for the per-group scalars of the step (computed once per
 * group by a small kernel instead of by every thread).                                                          */
int cgl_linear_wgrad_adam(int G, int rows, int in, int out, const float* dy, int64_t dy_gstride, const float* x,
                          int64_t x_gstride, float* params, float* adam_m, float* adam_v, int64_t ld,
                          const int32_t* step, const int32_t* ids, int64_t w_off, int64_t b_off, float lr, float beta1,
                          float beta2, float eps, void* adam_scratch, cgl_stream_t stream);

/* ---- convolutional LSGAN networks (model/lsgan.py:3-27 Generator, :73-99 Discriminator) -------------
 * The reference ships these classes without a call site; the engine runs them as implicit GEMMs over the grouped Linear
 * products above. Activations are [N images][H*W pixels][C channels] (NHWC; N = groups x batch), a 3x3 convolution
 * (padding 1, stride 1 or 2) of a group is   cgl_im2col3x3 -> cgl_linear_fwd(rows = B*OH*OW, in = C*9, out = Cout)   with
 * Conv2d.weight [Cout][Cin][3][3] read as the Linear weight [out][in] (im2col columns ordered (ci, kh, kw)), its backward
 * cgl_linear_wgrad_adam + cgl_linear_bwd_data -> cgl_col2im3x3. The host-side composition is cgl-gan_b200/conv.py.
 *   cgl_im2col3x3      col[n][oh][ow][ci*9+kh*3+kw] = x[n][oh*s-1+kh][ow*s-1+kw][ci]  (0 outside)
 *   cgl_col2im3x3      its transpose: dx[n][ih][iw][ci] = sum of the taps that read that pixel (gathered, fixed order)
 *   cgl_upsample2x     nn.Upsample(scale_factor=2) (nearest), H x W -> 2H x 2W;  _bwd: the sum of each 2 x 2 block
 *   cgl_channel_scale  x[n][pix][c] *= mask[n][c]: Dropout2d(0.25) with the mask injected (model/lsgan.py:79), also its backward
 *   cgl_nchw_to_nhwc / cgl_nhwc_to_nchw   out.view(B, 128, 8, 8) after l1 (:22-23) and out.view(B, -1) before adv_layer (:96-97)
 *   cgl_bn_forward / cgl_bn_backward   BatchNorm over the rows of [G][rows][F] (BatchNorm2d(C, 0.8) when rows are NHWC
 *                      pixels), parameters / running statistics in packed rows; backward takes gamma / beta's Adam step
 *   cgl_act_backward   dz = dy * act'(y)                                                                              */
int cgl_im2col3x3(int64_t N, int H, int W, int C, int stride, const float* x, float* col, cgl_stream_t stream);
int cgl_col2im3x3(int64_t N, int H, int W, int C, int stride, const float* dcol, float* dx, cgl_stream_t stream);
int cgl_upsample2x(int64_t N, int H, int W, int C, const float* x, float* y, cgl_stream_t stream);
int cgl_upsample2x_bwd(int64_t N, int H, int W, int C, const float* dy, float* dx, cgl_stream_t stream);
int cgl_channel_scale(int64_t N, int HW, int C, const float* mask, float* x, cgl_stream_t stream);
int cgl_nchw_to_nhwc(int64_t N, int C, int HW, const float* x, float* y, cgl_stream_t stream);
int cgl_nhwc_to_nchw(int64_t N, int C, int HW, const float* x, float* y, cgl_stream_t stream);
int cgl_bn_forward(int G, int rows, int F, const float* u, float* h, const float* params, int64_t ldp, const int32_t* ids,
                   int64_t gamma_off, int64_t beta_off, float* bn_stats, int64_t ld_stats, int64_t mean_off, int64_t var_off,
                   float* save_mean, float* save_invstd, float eps, float momentum, int train, int act, float slope,
                   cgl_stream_t stream);
int cgl_bn_backward(int G, int rows, int F, float* dz, const float* u, const float* save_mean, const float* save_invstd,
                    float* params, float* adam_m, float* adam_v, int64_t ldp, const int32_t* ids, int64_t gamma_off,
                    int64_t beta_off, const int32_t* step, float lr, float beta1, float beta2, float eps, cgl_stream_t stream);
int cgl_act_backward(int64_t n, const float* dy, const float* y, float* dz, int act, float slope, cgl_stream_t stream);
/* BatchNorm backward over two row segments normalised by two separate forward calls (net_d(real), net_d(fake): each pass
 * has its own batch statistics, their gradients meet in one optimizer step): dz / u are [G][rows0 + rows1][F]. */
int cgl_bn_backward_seg(int G, int rows0, int rows1, int F, float* dz, const float* u, const float* mean0, const float* invstd0,
                        const float* mean1, const float* invstd1, float* params, float* adam_m, float* adam_v, int64_t ldp,
                        const int32_t* ids, int64_t gamma_off, int64_t beta_off, const int32_t* step, float lr, float beta1,
                        float beta2, float eps, cgl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CGL_B200_H */
