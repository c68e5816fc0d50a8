"""Timing of single grouped Linear products for one CGL_TUNE setting (the ablation / probe bits of tc_tma.cuh):
    CGL_TUNE=<bits> python profiles/tma_probe.py fwd:1024:100:784 bwd:512:100:1024 ...      (kind:in:rows:out)"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgl_gan_b200 import abi  # noqa: E402

abi.require_device()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


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
for spec in sys.argv[1:]:
    kind, K, rows, out = spec.split(":")
    bench(kind, int(K), int(rows), int(out))
