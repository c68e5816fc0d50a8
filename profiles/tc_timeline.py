"""Per-CTA milestone timeline of the tcgen05 grouped GEMM (clock64 stamps, csrc/tc_gemm.cuh TC_STAMP).
    python profiles/tc_timeline.py [K] [rows] [out] [kind: fwd|wgrad]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgl_gan_b200 import abi  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 32
rows = int(sys.argv[2]) if len(sys.argv) > 2 else 100
out = int(sys.argv[3]) if len(sys.argv) > 3 else 128
kind = sys.argv[4] if len(sys.argv) > 4 else "fwd"
abi.require_device()
abi.check(abi.lib.cgl_set_gemm_mode(2))
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
G = int(sys.argv[5]) if len(sys.argv) > 5 else 148 * 8
ldp = (K * out + out + 31) // 32 * 32
prm = torch.randn(G, ldp, device="cuda") * 0.05
x = torch.randn(G, rows, K, device="cuda")
y = torch.empty(G, rows, out, device="cuda")
dy = torch.randn(G, rows, out, device="cuda")
grad = torch.zeros(G, ldp, device="cuda")
am, av = torch.zeros(G, ldp, device="cuda"), torch.zeros(G, ldp, device="cuda")
step = torch.ones(G, dtype=torch.int32, device="cuda")
scratch = torch.empty(G * 8, device="cuda")


def run():
    if kind == "fwd":
        abi.check(abi.lib.cgl_linear_fwd(G, rows, K, out, abi.ptr(x), rows * K, abi.ptr(prm), ldp, None, 0, K * out,
                                         abi.ACT_LRELU, 0.2, abi.ptr(y), rows * out, st()))
    elif kind == "wgrad":
        abi.check(abi.lib.cgl_linear_wgrad(G, rows, K, out, abi.ptr(dy), rows * out, abi.ptr(x), rows * K, abi.ptr(grad),
                                           ldp, None, 0, K * out, st()))
    else:   # "adam": the fused weight-gradient + Adam epilogue
        abi.check(abi.lib.cgl_linear_wgrad_adam(G, rows, K, out, abi.ptr(dy), rows * out, abi.ptr(x), rows * K,
                                                abi.ptr(prm), abi.ptr(am), abi.ptr(av), ldp, abi.ptr(step), None, 0,
                                                K * out, 2e-4, 0.5, 0.999, 1e-8, abi.ptr(scratch), st()))


run()
torch.cuda.synchronize()
buf = torch.zeros(G * 16, dtype=torch.int64, device="cuda")
abi.check(abi.lib.cgl_debug_set_timeline(abi.ptr(buf)))
run()
torch.cuda.synchronize()
abi.check(abi.lib.cgl_debug_set_timeline(None))
t = buf.view(G, 16).cpu().double()
names = ["entry", "setup done", "loads issued", "stage0 stored", "all stored", "acc complete", "epilogue done", "exit",
         "mma first full", "mma last commit"]
base = t[:, 0:1]
rel = (t[:, :10] - base) / 1.92e3     # us at ~1.92 GHz
for first in (slice(0, min(148, G)), slice(G // 2, min(G // 2 + 148, G))):
    r = rel[first]
    print("CTAs", first.start, "..", first.stop, " (median us since entry)")
    for i, n in enumerate(names):
        print(f"   {n:16s} {r[:, i].median():8.2f}")
