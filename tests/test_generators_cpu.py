"""StackedGenerator (batched over servers, torch ops) against the reference-style modules, on CPU:
forward, BatchNorm running statistics, and the split head / trunk backward of CGLGAN / Mix-G."""
import torch

from helpers import rel_err
from oracle import models as om


def _mods(kind, S, N):
    torch.manual_seed(3)
    if kind == "mnist_mix":
        return [om.MixGeneratorMNIST((1, 28, 28), N) for _ in range(S)]
    if kind == "2d_mix":
        return [om.Generator2DCGL((2,), N) for _ in range(S)]
    if kind == "mnist_plain":
        return [om.GeneratorMNIST((1, 28, 28)) for _ in range(S)]
    return [om.Generator2DMD((2,)) for _ in range(S)]


def test_stacked_generator_matches_modules():
    from cgl_gan_b200.generators import StackedGenerator
    for kind, shape, N in [("mnist_mix", (1, 28, 28), 3), ("2d_mix", (2,), 2), ("mnist_plain", (1, 28, 28), 0),
                           ("2d_plain", (2,), 0)]:
        S, B = 2, 16
        mods = _mods(kind, S, N)
        G = StackedGenerator(shape, S, N, device="cpu")
        G.load_modules(mods)
        z = torch.randn(S, B, 100)
        for rep in range(2):                       # two passes: running stats move twice
            out = G(z)
            for s in range(S):
                ref = mods[s](z[s])
                got = out[s].reshape(ref.shape) if N else out[s].reshape(ref.shape)
                assert rel_err(got, ref) < 1e-5, (kind, s, rel_err(got, ref))
        # split backward: heads receive d(sum loss), trunk d(sum w*loss)  (CGLGAN/2DMG/main.py:254-269)
        if N:
            w = torch.rand(S, N)
            tgt = torch.randn_like(out)
            G.zero_grad()
            out = G(z)
            G.trunk_scale["w"] = w.view(S, N, 1, 1)
            ((out - tgt) ** 2).mean(dim=(2, 3)).sum().backward()
            G.trunk_scale["w"] = None
            for s in range(S):
                m = mods[s]
                m.zero_grad()
                o = m(z[s]).reshape(N, B, -1)
                loss = ((o - tgt[s]) ** 2).mean(dim=(1, 2))
                m.model.requires_grad_(False)
                loss.sum().backward(retain_graph=True)
                m.model.requires_grad_(True)
                m.paths.requires_grad_(False)
                (w[s] * loss).sum().backward()
                m.paths.requires_grad_(True)
                gt = torch.cat([p.grad.reshape(-1) for p in m.model.parameters()])
                assert rel_err(G.trunk.params.grad[s, :G.P_trunk], gt) < 2e-5, (kind, "trunk")
                for i, path in enumerate(m.paths):
                    gh = torch.cat([p.grad.reshape(-1) for p in path.parameters()])
                    assert rel_err(G.heads.params.grad[s * N + i, :G.P_head], gh) < 2e-5, (kind, "head", i)
        # state round trip
        m2 = G.make_module()
        G.store_module(1, m2)
        for (k1, v1), (k2, v2) in zip(m2.state_dict().items(), mods[1].state_dict().items()):
            assert k1 == k2
            if v1.dim():
                assert rel_err(v1, v2) < 1e-5, k1
