"""Micro-benchmark of the dominant kernel: weight gradient + fused Adam (tcgen05 GEMM, EPI_ADAM) on one layer.
    python profiles/adam_bench.py [in] [rows] [out] [G] [reps]
Prints ms per launch and the algorithmic GB/s (24 B per parameter + the operands once). Run it under
`ncu --set full -k regex:tc_grouped -c 1` for the roofline.traffic capture."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgl_gan_b200 import abi  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 100
out = int(sys.argv[3]) if len(sys.argv) > 3 else 784
G = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
abi.require_device()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
ldp = (K * out + out + 31) // 32 * 32
prm = torch.randn(G, ldp, device="cuda") * 0.05
x = torch.randn(G, rows, K, device="cuda")
dy = torch.randn(G, rows, out, device="cuda") * 1e-3
am, av = torch.zeros(G, ldp, device="cuda"), torch.zeros(G, ldp, device="cuda")
step = torch.ones(G, dtype=torch.int32, device="cuda")
scratch = torch.empty(G * 8, device="cuda")


def run():
    abi.check(abi.lib.cgl_linear_wgrad_adam(G, rows, K, out, abi.ptr(dy), rows * out, abi.ptr(x), rows * K,
                                            abi.ptr(prm), abi.ptr(am), abi.ptr(av), ldp, abi.ptr(step), None, 0,
                                            K * out, 2e-4, 0.5, 0.999, 1e-8, abi.ptr(scratch), st()))


run()
torch.cuda.synchronize()
best = 1e9
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); run(); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
nbytes = G * (24.0 * (K * out + out) + 4.0 * rows * (K + out))
print(f"wgrad+adam in={K} rows={rows} out={out} G={G}: {best:.3f} ms  {nbytes / best / 1e6:.0f} GB/s algorithmic "
      f"({nbytes / 1e9:.2f} GB per launch)")
