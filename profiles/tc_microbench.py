"""Micro-benchmark of the tcgen05 grouped GEMM: per-CTA time against K (slope = mainloop cost per k-block,
intercept = prologue + epilogue), forward orientation, one 128 x bn tile per group.
    python profiles/tc_microbench.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgl_gan_b200 import abi  # noqa: E402

abi.require_device()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
dev = "cuda"


def time_fwd(G, rows, K, out, mode, reps=5):
    abi.check(abi.lib.cgl_set_gemm_mode(mode))
    ldp = K * out + out
    ldp = (ldp + 31) // 32 * 32
    prm = torch.randn(G, ldp, device=dev) * 0.05
    x = torch.randn(G, rows, K, device=dev)
    y = torch.empty(G, rows, out, device=dev)
    def run():
        abi.check(abi.lib.cgl_linear_fwd(G, rows, K, out, abi.ptr(x), rows * K, abi.ptr(prm), ldp, None, 0, K * out,
                                         abi.ACT_LRELU, 0.2, abi.ptr(y), rows * out, st()))
    run(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); run(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


waves = 8
G = 148 * waves
for rows in (100, 200):
    for K in (32, 64, 128, 256, 512, 1024):
        t = time_fwd(G, rows, K, 128, 2)
        f = time_fwd(G, rows, K, 128, 1)
        print(f"rows {rows:4d} K {K:5d}: tc {t*1e3/waves:8.1f} us per CTA-wave ({2*G*rows*K*128/t/1e9:7.1f} GFLOP/s)   ffma {f*1e3:8.1f} us total vs tc {t*1e3:8.1f}")
