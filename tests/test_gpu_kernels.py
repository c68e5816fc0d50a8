"""GPU parity of the C-ABI kernels against the oracle (torch CPU restatement of the reference).
Tolerances: fp32 parameters within 1e-5 relative (norm-wise), losses within 1e-4 (BASELINE.json north_star)."""
import random

import pytest
import torch

from helpers import D_IN, assert_grad_close, assert_params_close, assert_rows_close, make_batches, make_ds, rel_err, rel_l2, osteps

pytestmark = pytest.mark.gpu

PARAM_TOL = 1e-5
LOSS_TOL = 1e-4
LOSS_OF_ARCH = {0: 0, 1: 0, 2: 1, 3: 2}   # D_2D/D_MNIST1: BCE, D_MNIST2: CE, D_MNIST_LS: MSE


def _bank(lib, arch, nets, B, scale=1.0):
    from cgl_gan_b200.engine import ClientBank
    bank = ClientBank(arch, len(nets), B, loss_kind=LOSS_OF_ARCH[arch], d_loss_scale=scale)
    bank.load_modules(nets)
    return bank


@pytest.mark.parametrize("arch,scale", [(0, 1.0), (1, 1.0), (2, 0.5), (3, 1.0)])
@pytest.mark.parametrize("steps", [1, 3])
def test_d_step_and_g_loss_match_oracle(lib, arch, scale, steps):
    G, B = 5, 100
    kind = LOSS_OF_ARCH[arch]
    nets = make_ds(arch, G, seed=100 + arch)
    bank = _bank(lib, arch, nets, B, scale)
    loss = osteps.make_loss(kind)
    optis = [osteps.make_adam(n.parameters()) for n in nets]
    for it in range(steps):
        real, fake, xg = make_batches(arch, G, B, seed=7 * it + arch)
        n_real = torch.tensor([B, 41, 1, B, 77][:G])   # ragged last DataLoader batches
        real_pad = real.clone()
        for g in range(G):
            real_pad[g, n_real[g]:] = 0
        if it == 0:
            # generator-loss path alone, on identical parameters: tight fp32 agreement
            l_pre, dx_pre = bank.g_loss_raw(xg.cuda())
            for g in range(G):
                x = xg[g].clone().requires_grad_(True)
                l_ref = osteps.worker_g_loss(nets[g], loss, kind, x, B)
                l_ref.backward()
                assert abs(l_pre[g].item() - l_ref.item()) < 1e-5, (g, l_pre[g].item(), l_ref.item())
                assert_rows_close(dx_pre[g], x.grad, tag=g)
        d_gpu = bank.d_step(real_pad.cuda(), fake.cuda(), n_real=n_real)
        xg_dev = xg.cuda().requires_grad_(True)
        l_gpu = bank.g_loss(xg_dev)
        l_gpu.sum().backward()
        for g in range(G):
            d_ref = osteps.worker_d_step(nets[g], optis[g], loss, kind, real[g, :n_real[g]], fake[g], B, scale)
            x = xg[g].clone().requires_grad_(True)
            l_ref = osteps.worker_g_loss(nets[g], loss, kind, x, B)
            l_ref.backward()
            assert abs(d_gpu[g].item() - d_ref.item()) < LOSS_TOL, (it, g, d_gpu[g].item(), d_ref.item())
            assert abs(l_gpu[g].item() - l_ref.item()) < LOSS_TOL, (it, g)
            # after an Adam step the two D's differ at the ill-conditioned elements (see assert_grad_close)
            assert_grad_close(xg_dev.grad[g], x.grad, tag=(it, g))
    for g in range(G):
        # layer-wise (a small layer must not hide behind a large one)
        off = 0
        for p in nets[g].parameters():
            n = p.numel()
            assert_params_close(bank.rows()[g, off:off + n], p.reshape(-1), steps=steps, tag=(g, off), strict=(steps == 1))
            off += n
    assert bank.step.tolist() == [steps] * G


def test_d_step_indexed_clients_and_shared_fake(lib):
    """client_ids picks rows of the bank; fake_idx shares one server batch among its clients
    (MDGAN: every worker receives the same Xd, MDGAN/MNIST/mdgan.py:193-195)."""
    arch, C, B = 0, 6, 100
    nets = make_ds(arch, C, seed=5)
    bank = _bank(lib, arch, nets, B)
    before = bank.rows().clone()
    ids = torch.tensor([4, 1, 3])
    real, fake, xg = make_batches(arch, 3, B, seed=3, F=2)
    fake_idx = torch.tensor([1, 0, 1])
    bank.d_step(real.cuda(), fake.cuda(), fake_idx=fake_idx, client_ids=ids)
    loss = osteps.make_loss(0)
    for j, c in enumerate(ids.tolist()):
        opt = osteps.make_adam(nets[c].parameters())
        osteps.worker_d_step(nets[c], opt, loss, 0, real[j], fake[fake_idx[j]], B)
        ref = torch.cat([p.detach().reshape(-1) for p in nets[c].parameters()])
        assert_params_close(bank.rows()[c], ref, tag=c)
    for c in (0, 2, 5):   # untouched rows stay bit-identical
        assert torch.equal(bank.rows()[c], before[c])
    assert bank.step.tolist() == [0, 1, 0, 1, 1, 0]
    # shared Xg: the gradient is the per-server sum over its clients, weighted by the loss grads
    xg_dev = xg.cuda().requires_grad_(True)
    l = bank.g_loss(xg_dev, xg_idx=fake_idx, client_ids=ids)
    w = torch.tensor([0.2, 0.5, 0.3])
    (l * w.cuda()).sum().backward()
    x = xg.clone().requires_grad_(True)
    tot = 0
    for j, c in enumerate(ids.tolist()):
        tot = tot + w[j] * osteps.worker_g_loss(nets[c], loss, 0, x[fake_idx[j]], B)
    tot.backward()
    assert rel_l2(xg_dev.grad, x.grad) < 1e-3


@pytest.mark.parametrize("mode,tol", [(1, 2e-6), (2, 6e-6)])
def test_linear_blocks_against_torch(lib, mode, tol):
    """Grouped Linear fwd / bwd-data / wgrad for ragged shapes (not multiples of the tiles) on both GEMM kernels:
    mode 1 = FFMA (exact fp32 products), mode 2 = tcgen05 3xTF32 wherever the operands are float4-addressable
    (its fp32 accumulator truncates: ~2.4e-9 * K relative, see csrc/tc_gemm.cuh). float64 reference."""
    import ctypes as C
    from cgl_gan_b200 import abi
    abi.check(abi.lib.cgl_set_gemm_mode(mode))
    try:
        _linear_blocks(abi, tol)
    finally:
        abi.check(abi.lib.cgl_set_gemm_mode(0))


def _linear_blocks(abi, TOL):
    import ctypes as C
    torch.manual_seed(0)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for (G, rows, din, dout) in [(3, 100, 100, 32), (2, 200, 2, 128), (2, 37, 130, 257), (1, 300, 784, 512),
                                 (2, 100, 512, 1024), (2, 200, 512, 256), (1, 260, 132, 264), (2, 100, 1024, 784)]:
        ld = (din * dout + dout + 31) // 32 * 32
        prm = torch.randn(G, ld) * 0.1
        x = torch.randn(G, rows, din)
        dy = torch.randn(G, rows, dout)
        W = prm[:, :din * dout].view(G, dout, din)
        b = prm[:, din * dout:din * dout + dout]
        y_ref = torch.nn.functional.leaky_relu(torch.baddbmm(b.double().unsqueeze(1), x.double(), W.double().transpose(1, 2)), 0.2)
        prm_d, x_d, dy_d = prm.cuda(), x.cuda(), dy.cuda()
        y = torch.empty(G, rows, dout, device="cuda")
        abi.check(abi.lib.cgl_linear_fwd(G, rows, din, dout, abi.ptr(x_d), rows * din, abi.ptr(prm_d), ld, None, 0,
                                         din * dout, abi.ACT_LRELU, 0.2, abi.ptr(y), rows * dout, st))
        assert rel_err(y, y_ref) < TOL, (G, rows, din, dout, rel_err(y, y_ref))
        dx = torch.empty(G, rows, din, device="cuda")
        saved = torch.randn(G, rows, din)
        abi.check(abi.lib.cgl_linear_bwd_data(G, rows, din, dout, abi.ptr(dy_d), rows * dout, abi.ptr(prm_d), ld, None,
                                              0, abi.ptr(saved.cuda()), rows * din, abi.ACT_LRELU, 0.2, abi.ptr(dx),
                                              rows * din, st))
        dx_ref = torch.bmm(dy.double(), W.double()) * torch.where(saved > 0, 1.0, 0.2)
        assert rel_err(dx, dx_ref) < TOL, (G, rows, din, dout, rel_err(dx, dx_ref))
        grad = torch.zeros(G, ld, device="cuda")
        abi.check(abi.lib.cgl_linear_wgrad(G, rows, din, dout, abi.ptr(dy_d), rows * dout, abi.ptr(x_d), rows * din,
                                           abi.ptr(grad), ld, None, 0, din * dout, st))
        dW_ref = torch.bmm(dy.double().transpose(1, 2), x.double()).reshape(G, -1)
        db_ref = dy.double().sum(1)
        assert rel_err(grad[:, :din * dout], dW_ref) < TOL, (G, rows, din, dout)
        assert rel_err(grad[:, din * dout:din * dout + dout], db_ref) < 2e-6
        assert torch.all(grad[:, din * dout + dout:] == 0)


def test_adam_rows_matches_torch_adam(lib):
    from cgl_gan_b200.engine import adam_rows
    torch.manual_seed(1)
    R, n = 3, 1000
    p = torch.randn(R, n); p_ref = [p[r].clone().requires_grad_(True) for r in range(R)]
    opts = [torch.optim.Adam([q], lr=2e-4, betas=(0.5, 0.999)) for q in p_ref]
    pd, m, v = p.cuda(), torch.zeros(R, n, device="cuda"), torch.zeros(R, n, device="cuda")
    step = torch.zeros(R, dtype=torch.int32, device="cuda")
    for it in range(5):
        g = torch.randn(R, n) * (10.0 ** -it)
        adam_rows(pd, g.cuda(), m, v, step, 2e-4, 0.5, 0.999)
        for r in range(R):
            p_ref[r].grad = g[r].clone()
            opts[r].step()
    for r in range(R):
        assert rel_err(pd[r], p_ref[r]) < 1e-7
    assert step.tolist() == [5] * R


def test_mix_kernels_bit_exact(lib):
    """K3 accumulates in source order with separately rounded products and sums: bit-exact against the
    reference's dict loop (Cloud.run, CGLGAN/2DMG/main.py:126-133) restated by the oracle."""
    arch, C, B = 0, 7, 100
    nets = make_ds(arch, C, seed=11)
    bank = _bank(lib, arch, nets, B)
    A = torch.tensor([3., 1., 4., 1., 5., 9., 2.]); A /= A.sum()
    dicts = [osteps.copy_parameters(n) for n in nets]
    p = osteps.cloud_aggregate(dicts, A)
    ref = torch.cat([p[k].reshape(-1) for k in p])
    g = bank.weighted_sum(A)
    assert torch.equal(g[:bank.P].cpu(), ref)
    # segema mix back into every row
    before = bank.rows().cpu().clone()
    bank.broadcast(g, sigma=0.25)
    for c in range(C):
        self_p = {k: v for k, v in dicts[c].items()}
        mixed = osteps.segema_mix(self_p, p, 0.25)
        refc = torch.cat([mixed[k].reshape(-1) for k in mixed])
        assert torch.equal(bank.rows()[c].cpu(), refc), c
    bank.load_rows(before)
    # swap = permutation (MDGAN/MNIST/mdgan.py:158-164)
    rd = random.Random(100)
    order = osteps.mdgan_swap(list(range(C)), rd)
    M = torch.zeros(C, C)
    for idx in range(C):
        M[idx, order[idx]] = 1.0
    bank.mix(M)
    for idx in range(C):
        assert torch.equal(bank.rows()[idx].cpu(), before[order[idx]])
    # group mean over blocks of clients
    bank.load_rows(before)
    M = torch.zeros(C, C)
    groups = [[0, 1, 2], [3, 4], [5, 6]]
    for gr in groups:
        for i in gr:
            M[i, gr] = 1.0 / len(gr)
    bank.mix(M)
    for gr in groups:
        mean = osteps.group_mean([{"w": before[i].clone()} for i in gr])["w"]
        for i in gr:
            assert rel_err(bank.rows()[i], mean) < 1e-6
