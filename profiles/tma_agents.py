"""Where the agents of the TMA-fed kernel spend their main loop (library built with -DTCT_PROFILE, csrc/tc_tma.cuh):
    CGL_B200_LIB=.../libcgl_prof.so CGL_TUNE=<bits> python profiles/tma_agents.py [fwd|bwd] [in] [rows] [out] [G]
Medians over the CTAs (0, 0, g) of the clock64 stamps and of the cycles inside each agent's barrier waits."""
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
abi.require_device()
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
buf = torch.zeros(G * 32, dtype=torch.int64, device="cuda")
abi.check(abi.lib.cgl_debug_set_timeline(abi.ptr(buf)))
run()
torch.cuda.synchronize()
abi.check(abi.lib.cgl_debug_set_timeline(None))
t = buf.view(G, 32).cpu().double()
t = t[G // 4: 3 * G // 4]            # CTAs from the steady part of the grid
nkb = (K + 31) // 32 if kind == "fwd" else (out + 31) // 32
med = lambda v: float(v.median())
print(f"CGL_TUNE={os.environ.get('CGL_TUNE')} {kind} in={K} rows={rows} out={out}: {nkb} k-blocks; clocks (median over {t.shape[0]} CTAs)")
print(f"  set-up {med(t[:, 1] - t[:, 0]):8.0f}   entry -> MMA loop start {med(t[:, 2] - t[:, 0]):8.0f}   MMA loop {med(t[:, 3] - t[:, 2]):8.0f}"
      f" = {med(t[:, 3] - t[:, 2]) / nkb:6.0f} per k-block   last commit -> accumulators complete {med(t[:, 4] - t[:, 3]):8.0f}"
      f"   epilogue {med(t[:, 5] - t[:, 4]):8.0f}   whole CTA {med(t[:, 5] - t[:, 0]):8.0f}")
print(f"  MMA warp 0 issuing tcgen05.mma / tcgen05.commit: {med(t[:, 16]) / nkb:6.0f}  {med(t[:, 17]) / nkb:6.0f} per k-block of the tile"
      f"   waiting for the token (dual issue): {med(t[:, 18]) / nkb:6.0f}")
for name, w, tot in (("MMA warp   waits b_full / a_full", (6, 7), None), ("TMA thread waits b_empty / raw_empty", (8, 9), 15),
                     ("converter  waits raw_full / a_empty", (10, 11), 13), ("B warp     waits b_raw", (12,), 14)):
    ws = "  ".join(f"{med(t[:, i]) / nkb:6.0f}" for i in w)
    extra = f"   loop total {med(t[:, tot]) / nkb:6.0f}" if tot is not None else ""
    print(f"  {name:38s}: {ws} per k-block{extra}")
