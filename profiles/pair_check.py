"""Correctness (against float64) and timing of the grouped Linear forward / data-gradient products for one CGL_TUNE setting:
    CGL_TUNE=<bits> python profiles/pair_check.py
Used to compare the CTA-pair kernel (tc_pair.cuh, bits 256 / 512) with the one-CTA kernels on the shapes of a round."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgl_gan_b200 import abi  # noqa: E402

abi.require_device()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
torch.manual_seed(0)


def check(G, rows, din, dout):
    ld = (din * dout + dout + 31) // 32 * 32
    prm = torch.randn(G, ld) * 0.1
    x, dy, saved = torch.randn(G, rows, din), torch.randn(G, rows, dout), torch.randn(G, rows, din)
    W = prm[:, :din * dout].view(G, dout, din)
    b = prm[:, din * dout:din * dout + dout]
    y_ref = torch.nn.functional.leaky_relu(torch.baddbmm(b.double().unsqueeze(1), x.double(), W.double().transpose(1, 2)), 0.2)
    prm_d, x_d, dy_d = prm.cuda(), x.cuda(), dy.cuda()
    y = torch.empty(G, rows, dout, device="cuda")
    abi.check(abi.lib.cgl_linear_fwd(G, rows, din, dout, abi.ptr(x_d), rows * din, abi.ptr(prm_d), ld, None, 0, din * dout,
                                     abi.ACT_LRELU, 0.2, abi.ptr(y), rows * dout, st()))
    e_f = ((y.double().cpu() - y_ref).abs().max() / y_ref.abs().max()).item()
    # signed error along the result's own sign (a one-sided split / accumulation error shows here, a symmetric one averages out)
    b_f = (((y.double().cpu() - y_ref) * y_ref.sign()).mean() / y_ref.abs().mean()).item()
    dx = torch.empty(G, rows, din, device="cuda")
    abi.check(abi.lib.cgl_linear_bwd_data(G, rows, din, dout, abi.ptr(dy_d), rows * dout, abi.ptr(prm_d), ld, None, 0,
                                          abi.ptr(saved.cuda()), rows * din, abi.ACT_LRELU, 0.2, abi.ptr(dx), rows * din, st()))
    dx_ref = torch.bmm(dy.double(), W.double()) * torch.where(saved > 0, 1.0, 0.2)
    e_b = ((dx.double().cpu() - dx_ref).abs().max() / dx_ref.abs().max()).item()
    b_b = (((dx.double().cpu() - dx_ref) * dx_ref.sign()).mean() / dx_ref.abs().mean()).item()
    print(f"check G={G} rows={rows} in={din} out={dout}: fwd err {e_f:.2e} (bias {b_f:+.1e})  bwd err {e_b:.2e} (bias {b_b:+.1e})",
          flush=True)
    return max(e_f, e_b)


def bench(kind, K, rows, out, G=1024, reps=5):
    ldp = (K * out + out + 31) // 32 * 32
    prm = torch.randn(G, ldp, device="cuda") * 0.05
    x = torch.randn(G, rows, K, device="cuda")
    y = torch.empty(G, rows, out, device="cuda")
    dy = torch.randn(G, rows, out, device="cuda")
    dx = torch.empty(G, rows, K, device="cuda")

    def run():
        if kind == "fwd":
            abi.check(abi.lib.cgl_linear_fwd(G, rows, K, out, abi.ptr(x), rows * K, abi.ptr(prm), ldp, None, 0, K * out,
                                             abi.ACT_LRELU, 0.2, abi.ptr(y), rows * out, st()))
        else:
            abi.check(abi.lib.cgl_linear_bwd_data(G, rows, K, out, abi.ptr(dy), rows * out, abi.ptr(prm), ldp, None, 0,
                                                  abi.ptr(x), rows * K, abi.ACT_LRELU, 0.2, abi.ptr(dx), rows * K, st()))
    run()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    fl = 2.0 * G * rows * K * out
    print(f"bench {kind} in={K} rows={rows} out={out} G={G}: {best:.3f} ms  {fl / best / 1e9:.1f} TFLOP/s fp32-equivalent", flush=True)


print("CGL_TUNE =", os.environ.get("CGL_TUNE"))
worst = 0.0
for shape in [(2, 100, 1024, 784), (3, 200, 784, 512), (2, 100, 512, 1024), (2, 200, 512, 256), (1, 37, 132, 264), (2, 100, 256, 512)]:
    worst = max(worst, check(*shape))
print("worst error", worst)
if len(sys.argv) > 1 and sys.argv[1] == "bench":
    for kind, K, rows, out in [("fwd", 1024, 100, 784), ("fwd", 512, 100, 1024), ("fwd", 784, 200, 512), ("fwd", 512, 200, 256),
                               ("fwd", 784, 100, 512), ("bwd", 784, 100, 512), ("bwd", 512, 100, 1024), ("bwd", 1024, 100, 784),
                               ("bwd", 512, 200, 256)]:
        bench(kind, K, rows, out)
