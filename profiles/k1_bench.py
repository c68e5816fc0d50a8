"""Micro-benchmark of the fused 2DMG client step (csrc/client_fused.cuh): C clients, cgl_client_step in a loop.
    python profiles/k1_bench.py [clients] [reps]    -> ms per launch, client-steps/s, fp32 TFLOP/s (53.1 MFLOP per step)"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

ge.build()
from cgl_gan_b200 import abi  # noqa: E402
from cgl_gan_b200.engine import ClientBank  # noqa: E402

abi.lib.cgl_set_fused_client_step(int(os.environ.get("CGL_K1", "1")))
C = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
B = 100
torch.manual_seed(0)
bank = ClientBank(abi.ARCH_D_2D, C, B)
bank.load_rows(torch.randn(C, bank.P) * 0.05)
real = torch.tanh(torch.randn(C, B, 2, device="cuda"))
fake = torch.tanh(torch.randn(C // 2, B, 2, device="cuda") * 0.5)
xg = torch.tanh(torch.randn(C // 2, B, 2, device="cuda") * 0.5)
idx = (torch.arange(C) // 2).to(torch.int32).cuda()
for _ in range(3):
    bank.client_step(real, fake, xg, idx=idx)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(reps):
    bank.client_step(real, fake, xg, idx=idx)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / reps
print(f"k1 clients={C} ms={ms:.4f} client_steps_per_s={C / ms * 1e3:.0f} TFLOPps={C * 53.1e6 / ms / 1e9:.2f} "
      f"mode={abi.lib.cgl_get_gemm_mode()} K1={os.environ.get('CGL_K1', '1')}")
