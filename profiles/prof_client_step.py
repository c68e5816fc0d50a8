"""Profiling driver: the MNIST client step alone (cgl_d_step + cgl_g_loss over C clients), a few rounds.
    python profiles/prof_client_step.py [--clients 1024] [--rounds 3] [--mode 0|1|2]
Prints CUDA-event times per call; run it under `ncu --metrics gpu__time_duration.sum` for the launch list."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cgl_gan_b200 import abi  # noqa: E402
from cgl_gan_b200.engine import ClientBank  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clients", type=int, default=1024)
ap.add_argument("--rounds", type=int, default=3)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--arch", type=int, default=abi.ARCH_D_MNIST1)
a = ap.parse_args()

abi.require_device()
abi.check(abi.lib.cgl_set_gemm_mode(a.mode))
C, B = a.clients, 100
bank = ClientBank(a.arch, C, B, device="cuda:0")
d = bank.d
torch.manual_seed(0)
bank.params[:, :bank.P].normal_(0, 0.03)
real = torch.tanh(torch.randn(C, B, d, device="cuda"))
fake = torch.tanh(torch.randn(C, B, d, device="cuda") * 0.5)
xg = torch.tanh(torch.randn(C, B, d, device="cuda") * 0.5)
for r in range(a.rounds):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    bank.d_step(real, fake)
    e[1].record()
    bank.g_loss_raw(xg)
    e[2].record()
    torch.cuda.synchronize()
    print(f"round {r}: d_step {e[0].elapsed_time(e[1]):.3f} ms   g_loss {e[1].elapsed_time(e[2]):.3f} ms")
