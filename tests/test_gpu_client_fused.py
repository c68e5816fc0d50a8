"""K1, the fused shared-memory-resident client step (csrc/client_fused.cuh), against the oracle's Worker.train
(CGLGAN/2DMG/main.py:344-375) and against the layered kernels of the same library on identical inputs.
Bars: losses 1e-5 / 1e-4, parameters the bars of helpers.assert_params_close; fused vs layered on the same inputs before
any Adam step: dLoss/dXg within 5e-6 of its scale on EVERY element (K1: 3xTF32 products on mma.sync, ~5e-7 relative like the
tcgen05 kernels; layered FFMA mode: exact-fp32 FMA chains)."""
import pytest
import torch

from helpers import assert_grad_close, assert_params_close, make_batches, make_ds, osteps

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def fused_on(lib):
    """The fused kernel is a process-wide switch (cgl_set_fused_client_step): on for these tests, restored after."""
    was = lib.lib.cgl_get_fused_client_step()
    lib.check(lib.lib.cgl_set_fused_client_step(1))
    yield
    lib.check(lib.lib.cgl_set_fused_client_step(was))


def _bank(arch, nets, B, loss_kind=0):
    from cgl_gan_b200.engine import ClientBank
    bank = ClientBank(arch, len(nets), B, loss_kind=loss_kind)
    bank.load_modules(nets)
    return bank


def _launches(abi, fn):
    n0 = abi.launch_count()
    out = fn()
    return out, abi.launch_count() - n0


@pytest.mark.parametrize("B", [100, 37])
def test_client_step_is_one_launch_and_matches_oracle(lib, B):
    """cgl_client_step on the 2DMG discriminator: ONE kernel launch for D step + G loss + dLoss/dXg of all clients;
    ragged real batches (n_real = B, min(41, B), 1, 2; the empty batch is test_gpu_paths.py's), a fake / Xg batch shared by index."""
    abi = lib
    arch, G = 0, 6
    nets = make_ds(arch, G, seed=11)
    bank = _bank(arch, nets, B)
    loss = osteps.make_loss(0)
    optis = [osteps.make_adam(n.parameters()) for n in nets]
    idx = torch.tensor([0, 0, 1, 1, 2, 2])
    for it in range(2):
        real, fake, xg = make_batches(arch, G, B, seed=5 * it + B, F=3)
        n_real = torch.tensor([B, min(41, B), 1, 2, B, min(77, B)])
        real_pad = real.clone()
        for g in range(G):
            real_pad[g, n_real[g]:] = 0
        (d_gpu, l_gpu, dxg), n = _launches(abi, lambda: bank.client_step(real_pad.cuda(), fake.cuda(), xg.cuda(),
                                                                           n_real=n_real, idx=idx))
        assert n == 1, f"fused client step took {n} launches"
        for g in range(G):
            d_ref = osteps.worker_d_step(nets[g], optis[g], loss, 0, real[g, :n_real[g]], fake[idx[g]], B)
            x = xg[idx[g]].clone().requires_grad_(True)
            l_ref = osteps.worker_g_loss(nets[g], loss, 0, x, B)
            l_ref.backward()
            assert abs(d_gpu[g].item() - d_ref.item()) < 1e-4, (it, g, d_gpu[g].item(), d_ref.item())
            assert abs(l_gpu[g].item() - l_ref.item()) < 1e-4, (it, g, l_gpu[g].item(), l_ref.item())
            assert_grad_close(dxg[g], x.grad, tag=(it, g))
    for g in range(G):
        off = 0
        for p in nets[g].parameters():
            n = p.numel()
            assert_params_close(bank.rows()[g, off:off + n], p.reshape(-1), steps=2, tag=(g, off), strict=False)
            off += n
    assert bank.step.tolist() == [2] * G


def test_fused_and_layered_kernels_agree(lib):
    """The same inputs through K1 (automatic mode) and through the layered FFMA kernels (cgl_set_gemm_mode(1)):
    G loss and dLoss/dXg before any update agree to fp32 rounding on every element; after one Adam step the losses,
    Adam's m (linear in the gradient) and the parameters agree within the parity bars."""
    abi = lib
    arch, G, B = 0, 4, 100
    nets = make_ds(arch, G, seed=3)
    real, fake, xg = make_batches(arch, G, B, seed=9)
    n_real = torch.tensor([B, 63, B, 5])
    out = {}
    for mode in (0, 1):
        abi.check(abi.lib.cgl_set_gemm_mode(mode))
        try:
            bank = _bank(arch, nets, B)
            l0, dx0 = bank.g_loss_raw(xg.cuda())
            d = bank.d_step(real.cuda(), fake.cuda(), n_real=n_real)
            l1, dx1 = bank.g_loss_raw(xg.cuda())
            torch.cuda.synchronize()
            out[mode] = dict(l0=l0.cpu(), dx0=dx0.cpu(), d=d.cpu(), l1=l1.cpu(), dx1=dx1.cpu(), p=bank.rows().cpu(),
                             m=bank.adam_m[:, :bank.P].cpu(), v=bank.adam_v[:, :bank.P].cpu(), step=bank.step.cpu())
        finally:
            abi.check(abi.lib.cgl_set_gemm_mode(0))
    a, b = out[0], out[1]
    assert (a["l0"] - b["l0"]).abs().max() < 2e-6
    assert (a["dx0"] - b["dx0"]).abs().max() <= 5e-6 * b["dx0"].abs().max()
    assert (a["d"] - b["d"]).abs().max() < 2e-6
    # Adam's first moment after one step is (1 - beta1) * gradient: a direct, well-conditioned view of every gradient
    assert (a["m"] - b["m"]).abs().max() <= 5e-6 * b["m"].abs().max()
    assert (a["v"] - b["v"]).abs().max() <= 1e-5 * b["v"].abs().max()
    assert torch.equal(a["step"], b["step"])
    for g in range(G):
        assert_params_close(a["p"][g], b["p"][g], tag=g, strict=True)
    assert (a["l1"] - b["l1"]).abs().max() < 1e-4


def test_d_step_only_epochs_then_fused_tail(lib):
    """epoch == 2 (Worker.train's `for i in range(epoch)`): the first minibatch is a D step alone, the second one is
    fused with the G loss; MSE on the sigmoid output as the second supported loss."""
    abi = lib
    arch, G, B = 0, 3, 100
    nets = make_ds(arch, G, seed=21)
    bank = _bank(arch, nets, B, loss_kind=abi.LOSS_MSE)
    loss = osteps.make_loss(abi.LOSS_MSE)
    optis = [osteps.make_adam(n.parameters()) for n in nets]
    r0, f0, _ = make_batches(arch, G, B, seed=1)
    r1, f1, xg = make_batches(arch, G, B, seed=2)
    (_, n0) = _launches(abi, lambda: bank.d_step(r0.cuda(), f0.cuda()))
    (res, n1) = _launches(abi, lambda: bank.client_step(r1.cuda(), f1.cuda(), xg.cuda()))
    assert n0 == 1 and n1 == 1
    d_gpu, l_gpu, dxg = res
    for g in range(G):
        osteps.worker_d_step(nets[g], optis[g], loss, abi.LOSS_MSE, r0[g], f0[g], B)
        d_ref = osteps.worker_d_step(nets[g], optis[g], loss, abi.LOSS_MSE, r1[g], f1[g], B)
        x = xg[g].clone().requires_grad_(True)
        l_ref = osteps.worker_g_loss(nets[g], loss, abi.LOSS_MSE, x, B)
        l_ref.backward()
        assert abs(d_gpu[g].item() - d_ref.item()) < 1e-4 and abs(l_gpu[g].item() - l_ref.item()) < 1e-4
        assert_grad_close(dxg[g], x.grad, tag=g)
        ref = torch.cat([p.detach().reshape(-1) for p in nets[g].parameters()])
        assert_params_close(bank.rows()[g], ref, steps=2, tag=g, strict=False)
