"""Bring-up check of the tcgen05 grouped GEMM (not a pytest file): every Linear product through the C ABI in
FFMA mode and in tcgen05 mode against a float64 torch reference. Run on a B200:
    timeout 300 python tests/debug_tc.py
"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgl_gan_b200 import abi  # noqa: E402

dev = torch.device("cuda:0")
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def run(mode, G, rows, inn, out, seed=0):
    abi.check(abi.lib.cgl_set_gemm_mode(mode))
    g = torch.Generator().manual_seed(seed)
    ldp = inn * out + out + 12  # a row with slack; offsets multiple of 4
    params = torch.randn(G, ldp, generator=g).mul_(0.05).to(dev)
    x = torch.randn(G, rows, inn, generator=g).to(dev)
    dy = torch.randn(G, rows, out, generator=g).to(dev)
    w_off, b_off = 0, inn * out
    W = params[:, :inn * out].view(G, out, inn).double()
    b = params[:, b_off:b_off + out].double()
    res = {}
    # forward with LeakyReLU
    y = torch.empty(G, rows, out, device=dev)
    abi.check(abi.lib.cgl_linear_fwd(G, rows, inn, out, abi.ptr(x), rows * inn, abi.ptr(params), ldp, None, w_off,
                                     b_off, abi.ACT_LRELU, 0.2, abi.ptr(y), rows * out, st()))
    torch.cuda.synchronize()
    ref = torch.nn.functional.leaky_relu(torch.baddbmm(b.unsqueeze(1), x.double(), W.transpose(1, 2)), 0.2)
    res["fwd"] = rel(y, ref)
    # data gradient with saved activation (tanh derivative exercises the saved path)
    saved = torch.tanh(torch.randn(G, rows, inn, generator=g)).to(dev)
    dx = torch.empty(G, rows, inn, device=dev)
    abi.check(abi.lib.cgl_linear_bwd_data(G, rows, inn, out, abi.ptr(dy), rows * out, abi.ptr(params), ldp, None,
                                          w_off, abi.ptr(saved), rows * inn, abi.ACT_TANH, 0.2, abi.ptr(dx),
                                          rows * inn, st()))
    torch.cuda.synchronize()
    ref = torch.bmm(dy.double(), W) * (1 - saved.double() ** 2)
    res["bwd"] = rel(dx, ref)
    # weight gradient (stored)
    grad = torch.zeros(G, ldp, device=dev)
    abi.check(abi.lib.cgl_linear_wgrad(G, rows, inn, out, abi.ptr(dy), rows * out, abi.ptr(x), rows * inn,
                                       abi.ptr(grad), ldp, None, w_off, b_off, st()))
    torch.cuda.synchronize()
    refW = torch.bmm(dy.double().transpose(1, 2), x.double()).reshape(G, -1)
    refb = dy.double().sum(1)
    res["wgrad"] = rel(grad[:, :inn * out], refW)
    res["bgrad"] = rel(grad[:, b_off:b_off + out], refb)
    return res


if __name__ == "__main__":
    abi.require_device()
    shapes = [(2, 200, 784, 512), (3, 200, 512, 256), (2, 100, 512, 1024), (2, 100, 1024, 784), (2, 100, 100, 128),
              (1, 41, 256, 512), (2, 104, 128, 256)]
    bad = 0
    for (G, rows, inn, out) in shapes:
        r1 = run(1, G, rows, inn, out)
        r2 = run(2, G, rows, inn, out)
        print(f"G={G} rows={rows} in={inn} out={out}")
        for k in r1:
            flag = "" if r2[k] < 2e-6 else "   <-- TC off"
            bad += r2[k] >= 2e-6
            print(f"   {k:6s} ffma rel {r1[k]:.3e}   tc rel {r2[k]:.3e}{flag}")
    abi.check(abi.lib.cgl_set_gemm_mode(0))
    print("RESULT", "FAIL" if bad else "OK")
    sys.exit(1 if bad else 0)
