"""K3 -- aggregation over packed parameter rows, HBM roofline (bytes = 4 P (C_in + R_out)).
FL-GAN MNIST server step at 1024 clients: the uniform average of every client's D (P = 533,505) and G
(P = 1,510,032) rows (cgl_wsum), the load of the average back into every client (cgl_bcast_mix), and the
neighbour-D group mean as a mixing matrix (cgl_mix_csr, groups of 4).
    python profiles/mix_bench.py [--clients 1024]
Prints one JSON line per kernel; CUDA events on the launching stream, 3 warm-up + 10 timed launches, the
buffers (2.2 GB / 6.2 GB) are far larger than L2."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cgl_gan_b200 import abi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clients", type=int, default=1024)
a = ap.parse_args()
abi.require_device()
peak = 6650.0
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk)).get("hbm_gbs", peak)
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


Cn = a.clients
for name, P in (("D_mnist1", 533505), ("G_mnist", 1510032)):
    ld = (P + 31) // 32 * 32
    src = torch.randn(Cn, ld, device="cuda")
    out = torch.empty(ld, device="cuda")
    w = torch.full((Cn,), 1.0 / Cn, device="cuda")
    ms = timed(lambda: abi.check(abi.lib.cgl_wsum(Cn, ld, abi.ptr(w), None, abi.ptr(src), ld, abi.ptr(out), st())))
    by = 4.0 * ld * (Cn + 1)
    print(json.dumps({"kernel": "cgl_wsum", "rows": Cn, "P": P, "model": name, "ms": ms, "GBps": by / ms / 1e6,
                      "frac_of_measured_hbm": by / ms / 1e6 / peak}))
    ms = timed(lambda: abi.check(abi.lib.cgl_bcast_mix(Cn, ld, None, 0.0, abi.ptr(out), abi.ptr(src), ld, st())))
    by = 4.0 * ld * (Cn + 1)
    print(json.dumps({"kernel": "cgl_bcast_mix(sigma=0)", "rows": Cn, "P": P, "model": name, "ms": ms,
                      "GBps": by / ms / 1e6, "frac_of_measured_hbm": by / ms / 1e6 / peak}))
    ms = timed(lambda: abi.check(abi.lib.cgl_bcast_mix(Cn, ld, None, 0.5, abi.ptr(out), abi.ptr(src), ld, st())))
    by = 4.0 * ld * (2 * Cn + 1)
    print(json.dumps({"kernel": "cgl_bcast_mix(sigma=0.5)", "rows": Cn, "P": P, "model": name, "ms": ms,
                      "GBps": by / ms / 1e6, "frac_of_measured_hbm": by / ms / 1e6 / peak}))
    if name == "D_mnist1":
        dst = torch.empty_like(src)
        grp = 4
        row_ptr = torch.arange(0, Cn * grp + 1, grp, dtype=torch.int32, device="cuda")
        col = (torch.arange(Cn * grp, device="cuda") % grp + (torch.arange(Cn * grp, device="cuda") // (grp * grp)) * grp).to(torch.int32)
        vals = torch.full((Cn * grp,), 1.0 / grp, device="cuda")
        ms = timed(lambda: abi.check(abi.lib.cgl_mix_csr(Cn, ld, abi.ptr(row_ptr), abi.ptr(col), abi.ptr(vals), abi.ptr(src),
                                                         ld, abi.ptr(dst), ld, st())))
        by = 4.0 * ld * Cn * 2          # algorithmic: every row read once and written once (group reuse is on-chip)
        print(json.dumps({"kernel": "cgl_mix_csr(group mean of 4)", "rows": Cn, "P": P, "model": name, "ms": ms,
                          "GBps": by / ms / 1e6, "frac_of_measured_hbm": by / ms / 1e6 / peak}))
        del dst
    del src
