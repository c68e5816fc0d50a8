"""The convolutional LSGAN networks (model/lsgan.py) on the engine against the oracle's restated classes
(pinned to the reference's by tests/test_oracle_golden2.py): data-movement kernels against torch, the discriminator step
(MSE, two forward calls with their own BatchNorm2d statistics, injected Dropout2d masks, one Adam step), the generator
loss with dLoss/dXg, the generator forward / backward / Adam."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from helpers import assert_params_close, rel_err, rel_l2
from oracle import models as om
from oracle import steps as st

pytestmark = pytest.mark.gpu


def _s():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def test_conv_data_movement_kernels(lib):
    """im2col / col2im (stride 1 and 2), upsample and its backward, the NCHW <-> NHWC permutes against torch (exact: copies
    and short fixed-order sums)."""
    abi = lib
    g = torch.Generator().manual_seed(0)
    for (N, H, Cc, s) in [(3, 8, 5, 1), (2, 32, 1, 2), (2, 16, 16, 2), (1, 6, 7, 2)]:
        x = torch.randn(N, Cc, H, H, generator=g)
        x_nhwc = x.permute(0, 2, 3, 1).contiguous().cuda()
        OH = (H - 1) // s + 1
        col = torch.empty(N * OH * OH, Cc * 9, device="cuda")
        abi.check(abi.lib.cgl_im2col3x3(N, H, H, Cc, s, abi.ptr(x_nhwc), abi.ptr(col), _s()))
        ref = F.unfold(x, 3, padding=1, stride=s)                      # [N, C*9, L], rows ordered (ci, kh, kw)
        ref = ref.permute(0, 2, 1).reshape(N * OH * OH, Cc * 9)
        assert torch.equal(col.cpu(), ref)
        dcol = torch.randn(N * OH * OH, Cc * 9, generator=g)
        dx = torch.empty(N * H * H, Cc, device="cuda")
        abi.check(abi.lib.cgl_col2im3x3(N, H, H, Cc, s, abi.ptr(dcol.cuda()), abi.ptr(dx), _s()))
        dref = F.fold(dcol.reshape(N, OH * OH, Cc * 9).permute(0, 2, 1), (H, H), 3, padding=1, stride=s)
        dref = dref.permute(0, 2, 3, 1).reshape(N * H * H, Cc)
        assert rel_err(dx, dref) < 1e-6
    x = torch.randn(2, 3, 4, 4, generator=g)
    xn = x.permute(0, 2, 3, 1).contiguous().cuda()
    y = torch.empty(2 * 64, 3, device="cuda")
    abi.check(abi.lib.cgl_upsample2x(2, 4, 4, 3, abi.ptr(xn), abi.ptr(y), _s()))
    yref = F.interpolate(x, scale_factor=2).permute(0, 2, 3, 1).reshape(2 * 64, 3)
    assert torch.equal(y.cpu(), yref)
    dy = torch.randn(2, 3, 8, 8, generator=g)
    dyn = dy.permute(0, 2, 3, 1).contiguous().cuda()
    dx = torch.empty(2 * 16, 3, device="cuda")
    abi.check(abi.lib.cgl_upsample2x_bwd(2, 4, 4, 3, abi.ptr(dyn), abi.ptr(dx), _s()))
    dxref = F.avg_pool2d(dy, 2) * 4
    assert rel_err(dx, dxref.permute(0, 2, 3, 1).reshape(2 * 16, 3)) < 1e-6
    a = torch.randn(3, 5, 6, generator=g).cuda()
    b = torch.empty(3, 6, 5, device="cuda")
    abi.check(abi.lib.cgl_nchw_to_nhwc(3, 5, 6, abi.ptr(a), abi.ptr(b), _s()))
    assert torch.equal(b, a.permute(0, 2, 1).contiguous())
    c = torch.empty(3, 5, 6, device="cuda")
    abi.check(abi.lib.cgl_nhwc_to_nchw(3, 5, 6, abi.ptr(b), abi.ptr(c), _s()))
    assert torch.equal(c, a)


def _named(net):
    return {k: v.detach().clone() for k, v in net.state_dict().items() if v.dim()}


def test_conv_discriminator_step_and_generator_loss(lib):
    from cgl_gan_b200.conv import ConvDiscriminatorBank, sample_masks
    torch.manual_seed(3)
    G, B = 2, 8
    nets = [om.ConvDiscriminator(None) for _ in range(G)]
    bank = ConvDiscriminatorBank(G, B)
    bank.load_modules(nets)
    gen = torch.Generator().manual_seed(1)
    real = torch.tanh(torch.randn(G, B, 1, 32, 32, generator=gen))
    fake = torch.tanh(torch.randn(G, B, 1, 32, 32, generator=gen) * 0.5)
    xg = torch.tanh(torch.randn(G, B, 1, 32, 32, generator=gen) * 0.5)
    m_r, m_f, m_g = (sample_masks(G, B, generator=gen, device="cpu") for _ in range(3))
    cu = lambda ms: [m.cuda() for m in ms]
    d_loss = bank.d_step(real.reshape(G, B, 1024).cuda(), fake.reshape(G, B, 1024).cuda(), cu(m_r), cu(m_f))
    g_loss, dxg = bank.g_loss(xg.reshape(G, B, 1024).cuda(), cu(m_g))
    torch.cuda.synchronize()
    mse = st.make_loss(st.LOSS_MSE)
    for g in range(G):
        net = nets[g]
        opt = st.make_adam(net.parameters())
        opt.zero_grad()
        loss = mse(net(real[g], [m[g] for m in m_r]), torch.ones(B, 1)) + mse(net(fake[g], [m[g] for m in m_f]), torch.zeros(B, 1))
        loss.backward()
        opt.step()
        assert abs(d_loss[g].item() - loss.item()) < 1e-5, (g, d_loss[g].item(), loss.item())
        x = xg[g].clone().requires_grad_(True)
        gl = mse(net(x, [m[g] for m in m_g]), torch.ones(B, 1))
        gl.backward()
        assert abs(g_loss[g].item() - gl.item()) < 1e-4
        assert rel_l2(dxg[g].reshape(-1), x.grad.reshape(-1)) < 2e-3, (g, rel_l2(dxg[g].reshape(-1), x.grad.reshape(-1)))
        # parameters, tensor by tensor (a bias that feeds a BatchNorm has an identically zero gradient: Adam amplifies noise there)
        o = 0
        for name, p in net.named_parameters():
            n = p.numel()
            got = bank.flat_rows()[g, o:o + n]
            if name in ("model.3.bias", "model.7.bias", "model.11.bias"):
                assert (got.cpu() - p.detach().reshape(-1)).abs().max().item() <= 2.2 * 2e-4, name
            else:
                assert_params_close(got, p.reshape(-1), steps=1, tag=(g, name), strict=False, bulk=2e-5)
            o += n
        stats = torch.cat([b.reshape(-1) for nme, b in net.named_buffers() if "running" in nme])
        assert rel_err(bank.stats[g, :stats.numel()], stats) < 1e-4
    assert bank.step.tolist() == [1] * G


def test_conv_generator_forward_backward(lib):
    from cgl_gan_b200.conv import ConvGeneratorStack
    torch.manual_seed(4)
    S, B = 2, 4
    nets = [om.ConvGenerator(None) for _ in range(S)]
    gs = ConvGeneratorStack(S)
    gs.load_modules(nets)
    gen = torch.Generator().manual_seed(2)
    z_d, z_g = torch.randn(S, B, 100, generator=gen), torch.randn(S, B, 100, generator=gen)
    dy = torch.randn(S, B, 1024, generator=gen) * 0.01
    xd = gs(z_d.cuda())                       # the no_grad pass: only its BatchNorm side effect survives
    xg = gs(z_g.cuda())
    gs.backward_step(dy.cuda())
    torch.cuda.synchronize()
    for s in range(S):
        net = nets[s]
        opt = st.make_adam(net.parameters())
        with torch.no_grad():
            ref_d = net(z_d[s])
        ref_g = net(z_g[s])
        assert rel_err(xd[s], ref_d.reshape(B, 1024)) < 2e-5 and rel_err(xg[s], ref_g.reshape(B, 1024)) < 2e-5
        opt.zero_grad()
        (ref_g.reshape(B, 1024) * dy[s]).sum().backward()
        opt.step()
        o = 0
        for name, p in net.named_parameters():
            n = p.numel()
            got = gs.flat_rows()[s, o:o + n]
            if name in ("conv_blocks.1.bias", "conv_blocks.5.bias"):      # biases that feed a BatchNorm
                assert (got.cpu() - p.detach().reshape(-1)).abs().max().item() <= 2.2 * 2e-4, name
            else:
                assert_params_close(got, p.reshape(-1), steps=1, tag=(s, name), strict=False, bulk=2e-5)
            o += n
        stats = torch.cat([b.reshape(-1) for nme, b in net.named_buffers() if "running" in nme])
        assert rel_err(gs.stats[s, :stats.numel()], stats) < 1e-4
