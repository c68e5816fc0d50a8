"""Micro-benchmark of one grouped Linear product on the engine's GEMM kernels.
    python profiles/linear_bench.py <fwd|bwd> [in] [rows] [out] [G] [reps] [mode]
Prints ms per launch and fp32-equivalent TFLOP/s. Run under `ncu --set full -k regex:tc_grouped -c 1`."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgl_gan_b200 import abi  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "fwd"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 100
out = int(sys.argv[4]) if len(sys.argv) > 4 else 784
G = int(sys.argv[5]) if len(sys.argv) > 5 else 1024
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 5
mode = int(sys.argv[7]) if len(sys.argv) > 7 else 0
abi.require_device()
abi.check(abi.lib.cgl_set_gemm_mode(mode))
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
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
print(f"{kind} in={K} rows={rows} out={out} G={G}: {best:.3f} ms  {fl / best / 1e9:.1f} TFLOP/s fp32-equivalent")
