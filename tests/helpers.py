"""Shared helpers of the parity tests: seeded inputs, oracle-side modules, comparison metrics."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import models as om  # noqa: E402
from oracle import steps as osteps  # noqa: E402

D_CLASSES = {
    0: lambda: om.Discriminator2D(),
    1: lambda: om.DiscriminatorMNIST1((1, 28, 28)),
    2: lambda: om.DiscriminatorMNIST2((1, 28, 28)),
    3: lambda: om.DiscriminatorMNISTLS((1, 28, 28)),
}
D_IN = {0: 2, 1: 784, 2: 784, 3: 784}


def rel_err(a, b):
    """Norm-wise relative error max|a-b| / max|b| (the 1e-5 bar of BASELINE.json is read this way:
    relative to the tensor's scale, since element-wise ratios are meaningless next to zero)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def make_ds(arch, n, seed):
    g = torch.Generator().manual_seed(seed)
    nets = []
    for _ in range(n):
        torch.manual_seed(int(torch.randint(0, 2 ** 31 - 1, (1,), generator=g)))
        nets.append(D_CLASSES[arch]())
    return nets


def make_batches(arch, G, B, seed, F=None):
    g = torch.Generator().manual_seed(seed)
    d = D_IN[arch]
    real = torch.tanh(torch.randn(G, B, d, generator=g))
    fake = torch.tanh(torch.randn(F if F else G, B, d, generator=g) * 0.5)
    xg = torch.tanh(torch.randn(F if F else G, B, d, generator=g) * 0.5)
    return real, fake, xg
