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


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def max_abs(a, b):
    return (a.detach().double().cpu() - b.detach().double().cpu()).abs().max().item()


def frac_outside(a, b, rel=1e-5):
    """Fraction of elements with |a-b| > rel * max|b|."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs() > rel * b.abs().max()).double().mean().item()


def assert_params_close(a, b, lr=2e-4, steps=1, tag=""):
    """The parameter bar of BASELINE.json ("fp32 parameters within 1e-5 relative"), stated so that the
    reference passes it against itself:
      * >= 99.9 % of the elements of every parameter tensor within 1e-5 * max|ref|;
      * the rest bounded by the size of the Adam steps taken (a flipped direction is 2*lr per step);
      * ||a-b||_2 / ||ref||_2 < 1e-4 (no systematic error).
    Why not a plain max-norm: an fp32 Adam step is ill-conditioned at two kinds of elements -- gradients
    that cancel to below Adam's eps=1e-8 (update lr*g/(|g|+eps) has slope lr/eps = 2e4), and whole hidden
    units whose pre-activation lands within rounding of LeakyReLU's kink for some sample (derivative 1
    vs 0.2). Summation ORDER decides those; torch CPU with 1 vs 8 threads differs from itself by ~3e-5
    max-norm after one step (tests/test_oracle_golden.py::test_reference_self_noise)."""
    fo, e2, em = frac_outside(a, b), rel_l2(a, b), max_abs(a, b)
    assert fo <= 1e-3, (tag, "fraction outside 1e-5", fo)
    assert e2 < 1e-4, (tag, "rel_l2", e2)
    assert em <= 2.2 * lr * steps, (tag, "max_abs", em)


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
