"""StackedGenerator (cgl_mlp_forward / cgl_mlp_backward over the server axis) against the reference-style
modules run by torch on the CPU: forward, BatchNorm running statistics, the split head / trunk backward of
CGLGAN / Mix-G (CGLGAN/2DMG/main.py:254-269) and the fused Adam step -- on both GEMM kernels."""
import pytest
import torch

from helpers import assert_params_close, bn_fed_biases, max_abs, rel_err
from oracle import models as om
from oracle import steps as osteps

pytestmark = pytest.mark.gpu


def _mods(kind, S, N):
    torch.manual_seed(3)
    if kind == "mnist_mix":
        return [om.MixGeneratorMNIST((1, 28, 28), N) for _ in range(S)]
    if kind == "2d_mix":
        return [om.Generator2DCGL((2,), N) for _ in range(S)]
    if kind == "mnist_plain":
        return [om.GeneratorMNIST((1, 28, 28)) for _ in range(S)]
    return [om.Generator2DMD((2,)) for _ in range(S)]


@pytest.mark.parametrize("mode", [1, 2])   # 1: FFMA grouped GEMM, 2: tcgen05 3xTF32 grouped GEMM
@pytest.mark.parametrize("kind,shape,N", [("mnist_mix", (1, 28, 28), 3), ("2d_mix", (2,), 2),
                                          ("mnist_plain", (1, 28, 28), 0), ("2d_plain", (2,), 0)])
def test_stacked_generator_matches_modules(lib, mode, kind, shape, N):
    from cgl_gan_b200 import abi
    from cgl_gan_b200.generators import StackedGenerator
    abi.check(abi.lib.cgl_set_gemm_mode(mode))
    try:
        S, B = 2, 100
        mods = _mods(kind, S, N)
        optis = [osteps.make_adam(m.parameters()) for m in mods]
        G = StackedGenerator(shape, S, N)
        G.load_modules(mods)
        g = torch.Generator().manual_seed(11)
        for step in range(2):
            outs = []
            for rep in range(2):                       # Xd pass then Xg pass: running stats move twice
                z = torch.randn(S, B, 100, generator=g)
                out = G(z.cuda())
                refs = [mods[s](z[s]) for s in range(S)]
                for s in range(S):
                    assert rel_err(out[s].reshape(refs[s].shape), refs[s]) < 2e-5, (kind, step, rep, s)
            # split backward: heads receive d(sum loss), trunk d(sum w*loss); then Adam on everything
            w = torch.rand(S, max(N, 1), generator=g)
            tgt = torch.randn(out.shape, generator=g)
            dy = 2 * (out.cpu() - tgt) / (B * out.shape[-1])          # d/d out of mean((out - tgt)^2) per head
            if N:
                G.backward_step(dy.cuda(), trunk_w=w.cuda())
            else:
                G.backward_step((dy * w.view(S, 1, 1)).cuda())
            for s in range(S):
                m = mods[s]
                optis[s].zero_grad()
                o = refs[s].reshape(max(N, 1), B, -1)
                loss = ((o - tgt[s].reshape(o.shape)) ** 2).mean(dim=(1, 2))
                if N:
                    m.model.requires_grad_(False)
                    loss.sum().backward(retain_graph=True)
                    m.model.requires_grad_(True)
                    m.paths.requires_grad_(False)
                    (w[s] * loss).sum().backward()
                    m.paths.requires_grad_(True)
                else:
                    (w[s] * loss).sum().backward()
                optis[s].step()
        for s in range(S):
            got = G.make_module()
            G.store_module(s, got)
            skip = bn_fed_biases(mods[s])
            for (k1, v1), (k2, v2) in zip(got.state_dict().items(), mods[s].state_dict().items()):
                assert k1 == k2
                if not v1.dim():
                    continue
                if "running_mean" in k1 or k1 in skip:
                    assert max_abs(v1, v2) <= 2.2 * 2e-4 * 2, (kind, k1)
                elif "running_var" in k1:
                    assert rel_err(v1, v2) < 1e-4, (kind, k1)
                else:
                    assert_params_close(v1, v2, steps=2, tag=(kind, k1), strict=False, bulk=1e-4)
    finally:
        abi.check(abi.lib.cgl_set_gemm_mode(0))
