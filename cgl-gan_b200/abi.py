"""ctypes binding of include/cgl_b200.h. There is no CPU fallback: if lib/libcgl_b200.so is missing the
import of this module raises, and every compute call needs a CUDA device."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# CGL_B200_LIB: another build of the same library (kernel variants compared under profiles/)
LIB_PATH = os.environ.get("CGL_B200_LIB") or os.path.join(HERE, "lib", "libcgl_b200.so")

MAX_LAYERS = 8
ACT_NONE, ACT_LRELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3
LOSS_BCE, LOSS_CE, LOSS_MSE = 0, 1, 2
GEMM_AUTO, GEMM_FFMA, GEMM_TC = 0, 1, 2
(ARCH_D_2D, ARCH_D_MNIST1, ARCH_D_MNIST2, ARCH_D_MNIST_LS, ARCH_G_2D_MD, ARCH_G_MNIST,
 ARCH_G_2D_TRUNK, ARCH_G_2D_HEAD, ARCH_G_MNIST_TRUNK, ARCH_G_MNIST_HEAD) = range(10)


class MlpDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dims", C.c_int32 * (MAX_LAYERS + 1)),
                ("act", C.c_int32 * MAX_LAYERS), ("bn", C.c_int32 * MAX_LAYERS),
                ("bn_eps", C.c_float), ("bn_momentum", C.c_float), ("lrelu_slope", C.c_float)]


class MlpLayout(C.Structure):
    _fields_ = [("n_params", C.c_int64), ("w_off", C.c_int64 * MAX_LAYERS), ("b_off", C.c_int64 * MAX_LAYERS),
                ("bn_w_off", C.c_int64 * MAX_LAYERS), ("bn_b_off", C.c_int64 * MAX_LAYERS),
                ("n_bn_stats", C.c_int64), ("bn_mean_off", C.c_int64 * MAX_LAYERS),
                ("bn_var_off", C.c_int64 * MAX_LAYERS)]


class TrainCfg(C.Structure):
    _fields_ = [("loss_kind", C.c_int32), ("d_loss_scale", C.c_float), ("lr", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float)]


class CglError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "(nvcc, sm_100a). The engine has no CPU or PyTorch fallback.")

lib = C.CDLL(LIB_PATH)

_p = C.c_void_p
_i32, _i64, _f32, _sz = C.c_int, C.c_int64, C.c_float, C.c_size_t

lib.cgl_version.restype = C.c_char_p
lib.cgl_last_error.restype = C.c_char_p
lib.cgl_device_ok.restype = C.c_int
lib.cgl_launch_count.restype = C.c_longlong
lib.cgl_arch_describe.argtypes = [_i32, C.POINTER(MlpDesc)]
lib.cgl_mlp_layout_of.argtypes = [C.POINTER(MlpDesc), C.POINTER(MlpLayout)]
lib.cgl_d_step_workspace_bytes.argtypes = [C.POINTER(MlpDesc), _i32, _i32]
lib.cgl_d_step_workspace_bytes.restype = _sz
lib.cgl_g_loss_workspace_bytes.argtypes = [C.POINTER(MlpDesc), _i32, _i32]
lib.cgl_g_loss_workspace_bytes.restype = _sz
lib.cgl_d_step.argtypes = [C.POINTER(MlpDesc), _i32, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _i32,
                           C.POINTER(TrainCfg), _p, _p, _sz, _p]
lib.cgl_client_step.argtypes = [C.POINTER(MlpDesc), _i32, _p, _p, _p, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _i32,
                                C.POINTER(TrainCfg), _p, _p, _p, _p, _sz, _p]
lib.cgl_g_loss.argtypes = [C.POINTER(MlpDesc), _i32, _p, _i64, _p, _p, _p, _i32, _i32, _p, _p, _p, _sz, _p]
lib.cgl_dxg_reduce.argtypes = [_i32, _p, _p, _p, _p, _i64, _p, _p]
lib.cgl_adam_rows.argtypes = [_i32, _i64, _i64, _p, _p, _p, _p, _p, _f32, _f32, _f32, _f32, _p]
lib.cgl_mix_csr.argtypes = [_i32, _i64, _p, _p, _p, _p, _i64, _p, _i64, _p]
lib.cgl_wsum.argtypes = [_i32, _i64, _p, _p, _p, _i64, _p, _p]
lib.cgl_bcast_mix.argtypes = [_i32, _i64, _p, _f32, _p, _p, _i64, _p]
lib.cgl_wsum_div.argtypes = [_i32, _i64, _f32, _i32, _p, _p, _i64, _p, _p]
lib.cgl_comm_unique_id.argtypes = [_p]
lib.cgl_comm_init.argtypes = [_i32, _i32, _p, C.POINTER(_p)]
lib.cgl_comm_destroy.argtypes = [_p]
lib.cgl_allreduce_sum.argtypes = [_p, _p, _i64, _p]
lib.cgl_mix_allreduce.argtypes = [_p, _i32, _i64, _p, _p, _p, _i64, _p, _p]
lib.cgl_mlp_workspace_bytes.argtypes = [C.POINTER(MlpDesc), _i32, _i32]
lib.cgl_mlp_workspace_bytes.restype = _sz
lib.cgl_mlp_forward.argtypes = [C.POINTER(MlpDesc), _i32, _p, _i64, _p, _p, _i64, _i32, _p, _i64, _p, _i32, _p, _p, _sz, _p]
lib.cgl_mlp_backward.argtypes = [C.POINTER(MlpDesc), _i32, _p, _p, _p, _i64, _p, _p, C.POINTER(TrainCfg), _p, _i64, _p,
                                 _i32, _p, _p, _p, _p, _sz, _p]
lib.cgl_fl_step_workspace_bytes.argtypes = [C.POINTER(MlpDesc), C.POINTER(MlpDesc), _i32, _i32]
lib.cgl_fl_step_workspace_bytes.restype = _sz
lib.cgl_fl_step.argtypes = [C.POINTER(MlpDesc), C.POINTER(MlpDesc), _i32, _p, _p, _p, _i64, _p, _p, _i64, _p, _p, _p, _i64, _p,
                            _p, _p, _p, _p, _p, _i32, C.POINTER(TrainCfg), C.POINTER(TrainCfg), _p, _p, _p, _sz, _p]
lib.cgl_im2col3x3.argtypes = [_i64, _i32, _i32, _i32, _i32, _p, _p, _p]
lib.cgl_col2im3x3.argtypes = [_i64, _i32, _i32, _i32, _i32, _p, _p, _p]
lib.cgl_upsample2x.argtypes = [_i64, _i32, _i32, _i32, _p, _p, _p]
lib.cgl_upsample2x_bwd.argtypes = [_i64, _i32, _i32, _i32, _p, _p, _p]
lib.cgl_channel_scale.argtypes = [_i64, _i32, _i32, _p, _p, _p]
lib.cgl_nchw_to_nhwc.argtypes = [_i64, _i32, _i32, _p, _p, _p]
lib.cgl_nhwc_to_nchw.argtypes = [_i64, _i32, _i32, _p, _p, _p]
lib.cgl_bn_forward.argtypes = [_i32, _i32, _i32, _p, _p, _p, _i64, _p, _i64, _i64, _p, _i64, _i64, _i64, _p, _p, _f32, _f32,
                               _i32, _i32, _f32, _p]
lib.cgl_bn_backward.argtypes = [_i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _i64, _p, _i64, _i64, _p, _f32, _f32, _f32, _f32, _p]
lib.cgl_act_backward.argtypes = [_i64, _p, _p, _p, _i32, _f32, _p]
lib.cgl_bn_backward_seg.argtypes = [_i32, _i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i64, _p, _i64, _i64, _p,
                                    _f32, _f32, _f32, _f32, _p]
lib.cgl_profile_enable.argtypes = [_i32]
lib.cgl_profile_tag_name.argtypes = [_i32]
lib.cgl_profile_tag_name.restype = C.c_char_p
lib.cgl_profile_summary.argtypes = [_i32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(C.c_longlong)]
lib.cgl_linear_wgrad_adam.argtypes = [_i32, _i32, _i32, _i32, _p, _i64, _p, _i64, _p, _p, _p, _i64, _p, _p, _i64, _i64,
                                      _f32, _f32, _f32, _f32, _p, _p]
lib.cgl_gather_rows.argtypes = [_i64, _i32, _p, _i64, _p, _p, _p]
lib.cgl_hist2d.argtypes = [_i64, _p, _i64, _p, _p]
lib.cgl_kl_score_2d.argtypes = [_i64, _p, _i64, _p, _p, _p, _p]
lib.cgl_set_gemm_mode.argtypes = [_i32]
lib.cgl_set_fused_client_step.argtypes = [_i32]
lib.cgl_set_fused_client_step.restype = C.c_int
lib.cgl_get_fused_client_step.restype = C.c_int
lib.cgl_debug_set_timeline.argtypes = [_p]
lib.cgl_get_gemm_mode.restype = C.c_int
lib.cgl_linear_fwd.argtypes = [_i32, _i32, _i32, _i32, _p, _i64, _p, _i64, _p, _i64, _i64, _i32, _f32, _p, _i64, _p]
lib.cgl_linear_bwd_data.argtypes = [_i32, _i32, _i32, _i32, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _i32, _f32,
                                    _p, _i64, _p]
lib.cgl_linear_wgrad.argtypes = [_i32, _i32, _i32, _i32, _p, _i64, _p, _i64, _p, _i64, _p, _i64, _i64, _p]

for _name in ("cgl_arch_describe", "cgl_mlp_layout_of", "cgl_d_step", "cgl_g_loss", "cgl_client_step", "cgl_dxg_reduce",
              "cgl_adam_rows", "cgl_mix_csr", "cgl_wsum", "cgl_bcast_mix", "cgl_wsum_div", "cgl_fl_step", "cgl_im2col3x3",
              "cgl_col2im3x3", "cgl_upsample2x", "cgl_upsample2x_bwd", "cgl_channel_scale", "cgl_nchw_to_nhwc", "cgl_nhwc_to_nchw",
              "cgl_bn_forward", "cgl_bn_backward", "cgl_act_backward", "cgl_bn_backward_seg", "cgl_comm_unique_id",
              "cgl_comm_init", "cgl_comm_destroy", "cgl_allreduce_sum", "cgl_mix_allreduce",
              "cgl_linear_fwd", "cgl_linear_bwd_data", "cgl_linear_wgrad", "cgl_set_gemm_mode", "cgl_mlp_forward", "cgl_mlp_backward", "cgl_profile_enable",
              "cgl_profile_summary", "cgl_debug_set_timeline", "cgl_linear_wgrad_adam", "cgl_gather_rows", "cgl_hist2d",
              "cgl_kl_score_2d"):
    getattr(lib, _name).restype = C.c_int


def check(rc):
    if rc != 0:
        raise CglError(f"cgl_b200 error {rc}: {lib.cgl_last_error().decode()}")


def version():
    return lib.cgl_version().decode()


PROF_NUM_TAGS = 14


def profile_enable(on=True):
    check(lib.cgl_profile_enable(1 if on else 0))


def profile_summary():
    """{kernel class: dict(ms, bytes, flops, launches)} since the last profile_enable (synchronises)."""
    out = {}
    for tag in range(PROF_NUM_TAGS):
        ms, by, fl, n = C.c_double(), C.c_double(), C.c_double(), C.c_longlong()
        check(lib.cgl_profile_summary(tag, C.byref(ms), C.byref(by), C.byref(fl), C.byref(n)))
        if n.value:
            out[lib.cgl_profile_tag_name(tag).decode()] = dict(ms=ms.value, bytes=by.value, flops=fl.value,
                                                               launches=n.value)
    return out


def launch_count():
    return int(lib.cgl_launch_count())


def device_ok():
    return bool(lib.cgl_device_ok())


def require_device():
    if not device_ok():
        raise CglError("no sm_100 CUDA device visible: the cgl_b200 engine has no CPU fallback")


def arch_describe(arch_id):
    d = MlpDesc()
    check(lib.cgl_arch_describe(arch_id, C.byref(d)))
    return d


def make_desc(dims, acts, bn=None, bn_eps=0.8, bn_momentum=0.1, slope=0.2):
    d = MlpDesc()
    n = len(dims) - 1
    assert 1 <= n <= MAX_LAYERS and len(acts) == n
    d.n_layers = n
    for i, v in enumerate(dims):
        d.dims[i] = int(v)
    for i in range(n):
        d.act[i] = int(acts[i])
        d.bn[i] = int(bn[i]) if bn is not None else 0
    d.bn_eps, d.bn_momentum, d.lrelu_slope = bn_eps, bn_momentum, slope
    return d


def layout_of(desc):
    lay = MlpLayout()
    check(lib.cgl_mlp_layout_of(C.byref(desc), C.byref(lay)))
    return lay


def ptr(t):
    """Device (or host) address of a torch tensor / None -> NULL."""
    return None if t is None else C.c_void_p(t.data_ptr())
