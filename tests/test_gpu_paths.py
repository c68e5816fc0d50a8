"""Shipped paths that round 1 left without a test ("what's weak" 1-2): the fused weight-gradient + Adam
kernel checked where it is well conditioned, adam_update_fast against the IEEE sequence, the as-written Cloud, capgan's
epoch-based cloud period, eval-mode generator snapshots, an empty real batch, FLGAN's full local passes."""
import ctypes as C

import pytest
import torch

from helpers import assert_params_close, make_batches, make_ds, osteps
from oracle import models as om
from oracle.rounds import OracleMD

pytestmark = pytest.mark.gpu


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---- (b) the FUSED path in the regime where Adam is linear in the gradient ------------------------------------------
@pytest.mark.parametrize("mode,tol", [(1, 3e-6), (2, 1e-5)], ids=["ffma", "tcgen05"])
@pytest.mark.parametrize("G,rows,din,dout", [(2, 200, 784, 512), (3, 100, 512, 1024), (2, 200, 512, 256),
                                             (2, 100, 1024, 784), (1, 200, 132, 260)])
def test_fused_wgrad_adam_is_exact_where_adam_is_linear(lib, mode, tol, G, rows, din, dout):
    """cgl_linear_wgrad_adam with eps >> |g|: the first Adam step is p -= lr * g / (|g| + eps), smooth in g, so the
    parameter delta is a conditioning-free read-out of the gradient the fused epilogue computed. Starting from p = 0 the
    delta is the stored value itself. Against float64, EVERY element within `tol` of the tensor's scale."""
    abi = lib
    abi.check(abi.lib.cgl_set_gemm_mode(mode))
    try:
        torch.manual_seed(din + dout)
        ld = (din * dout + dout + 31) // 32 * 32
        x = torch.randn(G, rows, din)
        dy = torch.randn(G, rows, dout) / rows
        p = torch.zeros(G, ld, device="cuda")
        m, v = torch.zeros_like(p), torch.zeros_like(p)
        step = torch.ones(G, dtype=torch.int32, device="cuda")
        scratch = torch.empty(32 * G, dtype=torch.uint8, device="cuda")
        lr, eps = 1.0e4, 1.0e4
        xd, dyd = x.cuda(), dy.cuda()
        abi.check(abi.lib.cgl_linear_wgrad_adam(G, rows, din, dout, abi.ptr(dyd), rows * dout, abi.ptr(xd), rows * din,
                                                abi.ptr(p), abi.ptr(m), abi.ptr(v), ld, abi.ptr(step), None, 0, din * dout,
                                                lr, 0.5, 0.999, eps, abi.ptr(scratch), _st()))
        torch.cuda.synchronize()
        gW = torch.bmm(dy.double().transpose(1, 2), x.double()).reshape(G, -1)
        gb = dy.double().sum(1)
        for g64, got, what in ((gW, p[:, :din * dout], "W"), (gb, p[:, din * dout:din * dout + dout], "b")):
            ref = -lr * g64 / (g64.abs() + eps)
            err = (got.double().cpu() - ref).abs().max().item() / ref.abs().max().item()
            assert err <= tol, (what, G, rows, din, dout, err)
            # the moments: m = (1 - b1) g, v = (1 - b2) g^2
            mm = m[:, :din * dout] if what == "W" else m[:, din * dout:din * dout + dout]
            assert ((mm.double().cpu() - 0.5 * g64).abs().max() / (0.5 * g64).abs().max()).item() <= tol
        assert torch.all(p[:, din * dout + dout:] == 0)
    finally:
        abi.check(abi.lib.cgl_set_gemm_mode(0))


def test_adam_update_fast_against_the_ieee_sequence(lib):
    """The tcgen05 epilogue's adam_update_fast (MUFU sqrt / rcp) against the FFMA kernel's IEEE adam_update on inputs whose
    gradient is EXACT in both GEMMs (small integers times a power of two: every tf32 split and every partial sum is exact),
    so any difference is the update arithmetic alone: m and v bit-identical, p within 1e-6 of the step it took."""
    abi = lib
    G, rows, din, dout = 2, 100, 512, 256
    ld = (din * dout + dout + 31) // 32 * 32
    gen = torch.Generator().manual_seed(3)
    res = {}
    for scale in (1.0, 2.0 ** -12, 2.0 ** -24):
        x = torch.randint(-3, 4, (G, rows, din), generator=gen).float()
        dy = torch.randint(-2, 3, (G, rows, dout), generator=gen).float() * scale
        p0 = torch.randn(G, ld, generator=gen) * 0.05
        m0 = torch.randn(G, ld, generator=gen) * 1e-3 * scale
        v0 = (torch.randn(G, ld, generator=gen) * 1e-3 * scale) ** 2
        for mode in (1, 2):
            abi.check(abi.lib.cgl_set_gemm_mode(mode))
            p, m, v = p0.cuda(), m0.cuda(), v0.cuda()
            step = torch.full((G,), 7, dtype=torch.int32, device="cuda")
            xd, dyd = x.cuda(), dy.cuda()
            abi.check(abi.lib.cgl_linear_wgrad_adam(G, rows, din, dout, abi.ptr(dyd), rows * dout, abi.ptr(xd), rows * din,
                                                    abi.ptr(p), abi.ptr(m), abi.ptr(v), ld, abi.ptr(step), None, 0,
                                                    din * dout, 2e-4, 0.5, 0.999, 1e-8, None, _st()))
            torch.cuda.synchronize()
            res[mode] = (p.cpu(), m.cpu(), v.cpu())
        abi.check(abi.lib.cgl_set_gemm_mode(0))
        (p1, m1, v1), (p2, m2, v2) = res[1], res[2]
        n = din * dout
        assert torch.equal(m1[:, :n], m2[:, :n]) and torch.equal(v1[:, :n], v2[:, :n]), scale
        step_taken = (p1[:, :n] - p0[:, :n]).abs()
        diff = (p1[:, :n] - p2[:, :n]).abs()
        # <= 1e-6 of the step taken, plus one rounding of p itself (the two versions round the final sum separately)
        bound = 1e-6 * step_taken + 1.2e-7 * p0[:, :n].abs()
        assert bool((diff <= bound).all()), (scale, (diff - bound).max().item())
        assert step_taken.max().item() > 1e-5          # the step is not degenerate


# ---- (c) paths -----------------------------------------------------------------------------------------------------
def _round_inputs(C_, S, B, d, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.tanh(torch.randn(1, C_, B, d, generator=g)), torch.full((1, C_), B, dtype=torch.int32),
            torch.randn(S, B, 100, generator=g), torch.randn(S, B, 100, generator=g))


def test_cloud_as_written_is_a_no_op(lib):
    """cloud_mode="as_written": net_g.load_state_dict(recv_p, strict=False) with trunk-relative keys loads nothing
    (SURVEY 3.5.2), so the servers' trunks never meet; the engine follows and matches the oracle in that mode."""
    from cgl_gan_b200.sim import Knobs, MDStyleSim
    torch.manual_seed(3)
    W, S, B, shape, d = 4, 2, 100, (2,), 2
    orc = OracleMD("cglgan", W, S, B, shape, iid=1, cloud_mode="as_written")
    sim = MDStyleSim("cglgan", Knobs(num_workers=W, num_servers=S, batch_size=B, iid=1, img_shape=shape, cloud_mode="as_written"))
    sim.load(orc.net_g, orc.net_d)
    t0 = sim.G.trunk.params.clone()
    assert not torch.equal(t0[0], t0[1])
    sim.cloud_aggregate()
    assert torch.equal(sim.G.trunk.params, t0)
    for r in range(2):
        real, n_real, z_d, z_g = _round_inputs(W, S, B, d, 70 + r)
        l_ref = orc.round(real, n_real, z_d, z_g)
        l = sim.round(real.cuda(), n_real.cuda(), z_d.cuda(), z_g.cuda())
        assert (l.cpu() - l_ref).abs().max() < 1e-4
    for s in range(S):
        ref = torch.cat([p.detach().reshape(-1) for p in orc.net_g[s].model.parameters()])
        assert_params_close(sim.G.trunk.params[s, :ref.numel()], ref, steps=2, tag=("trunk", s), strict=False, bulk=1e-4)
    assert not torch.equal(sim.G.trunk.params[0], sim.G.trunk.params[1])


def test_capgan_cloud_period_counts_epochs(lib):
    """capgan.py:169: the cloud exchange fires when t % (data_len * cloud_epoch / batch_size) == 0 with t counting DOWN.
    Servers with 200 samples each and batch 100 meet every 2 rounds; servers whose periods disagree would deadlock the
    reference's rendezvous and are refused."""
    from cgl_gan_b200.sim import Knobs, MDStyleSim
    torch.manual_seed(5)
    W, S, B, shape, d = 4, 2, 100, (1, 28, 28), 784
    sizes = [100] * W
    orc = OracleMD("capgan", W, S, B, shape, part_sizes=sizes, num_communication=4)
    sim = MDStyleSim("capgan", Knobs(num_workers=W, num_servers=S, batch_size=B, img_shape=shape, num_communication=4),
                     part_sizes=sizes)
    sim.load(orc.net_g, orc.net_d)
    fired = []
    inner = sim.cloud_aggregate
    sim.cloud_aggregate = lambda: (fired.append(sim.t), inner())
    for r in range(3):
        real, n_real, z_d, z_g = _round_inputs(W, S, B, d, 80 + r)
        l_ref = orc.round(real, n_real, z_d, z_g)
        l = sim.round(real.cuda(), n_real.cuda(), z_d.cuda(), z_g.cuda())
        assert (l.cpu() - l_ref).abs().max() < 1e-4
    assert fired == [0, 2]                       # t = 4 and t = 2
    for s in range(S):
        ref = torch.cat([p.detach().reshape(-1) for p in orc.net_g[s].parameters()])
        assert_params_close(sim.G.trunk.params[s, :ref.numel()], ref, steps=3, tag=("G", s), strict=False, bulk=1e-4)
    bad = MDStyleSim("capgan", Knobs(num_workers=W, num_servers=S, batch_size=B, img_shape=shape, num_communication=6),
                     part_sizes=[100, 100, 150, 150])
    bad.t = 2                                     # t = 4: period 2 is due, period 3 is not
    with pytest.raises(ValueError):
        bad._cloud_due()


@pytest.mark.parametrize("heads", [0, 2])
def test_eval_forward_uses_running_statistics(lib, heads):
    """G.eval() snapshots (plot_2d, CGLGAN/2DMG/main.py:217-223): cgl_mlp_forward(train=0) normalises with the running
    statistics two training passes accumulated and leaves them untouched."""
    from cgl_gan_b200.generators import StackedGenerator
    torch.manual_seed(9)
    S, B = 2, 100
    mods = [om.MixGeneratorMNIST((1, 28, 28), heads) if heads else om.GeneratorMNIST((1, 28, 28)) for _ in range(S)]
    G = StackedGenerator((1, 28, 28), S, heads)
    G.load_modules(mods)
    g = torch.Generator().manual_seed(2)
    for _ in range(2):
        z = torch.randn(S, B, 100, generator=g)
        G(z.cuda())
        for s in range(S):
            mods[s](z[s])
    z = torch.randn(S, 16, 100, generator=g)
    before = G.trunk.stats.clone()
    G.eval()
    y = G(z.cuda()).cpu()
    G.train()
    assert torch.equal(G.trunk.stats, before)
    for s in range(S):
        mods[s].eval()
        with torch.no_grad():
            ref = mods[s](z[s])
        mods[s].train()
        got = y[s].reshape(ref.shape)
        assert (got - ref).abs().max().item() < 2e-5, (heads, s, (got - ref).abs().max().item())


def test_empty_real_batch_contributes_nothing(lib):
    """n_real = 0 (never produced by a DataLoader; torch's mean over an empty batch would be NaN): the real term is
    dropped, the step equals the fake-only step."""
    from cgl_gan_b200.engine import ClientBank
    arch, G, B = 0, 2, 100
    nets = make_ds(arch, G, seed=31)
    bank = ClientBank(arch, G, B)
    bank.load_modules(nets)
    real, fake, _ = make_batches(arch, G, B, seed=4)
    n_real = torch.tensor([0, B], dtype=torch.int32)
    real[0] = 0
    dl = bank.d_step(real.cuda(), fake.cuda(), n_real=n_real)
    loss = osteps.make_loss(0)
    opt = osteps.make_adam(nets[0].parameters())
    opt.zero_grad()
    l0 = loss(nets[0](fake[0]), torch.zeros(B, 1))
    l0.backward()
    opt.step()
    assert torch.isfinite(dl).all() and abs(dl[0].item() - l0.item()) < 1e-5
    ref = torch.cat([p.detach().reshape(-1) for p in nets[0].parameters()])
    assert_params_close(bank.rows()[0], ref, tag="empty real batch")


@pytest.mark.parametrize("d_share,algo", [("swap", "mdgan"), ("group_mean", "acgan")])
def test_discriminator_sharing_inside_a_round(lib, d_share, algo):
    """Knob E > 0 (README.md:26): the share happens inside MDStyleSim.round at the rounds the reference's test selects,
    swap = exact row copies, group mean = receive_parameter's sum-then-divide, bit-exact on the shared parameters."""
    from cgl_gan_b200.sim import Knobs, MDStyleSim
    torch.manual_seed(13)
    shape, d, B = (2,), 2, 100
    W, S = (4, 1) if algo == "mdgan" else (6, 2)
    kw = dict(E=2, d_share=d_share, num_communication=4)
    orc = OracleMD(algo, W, S, B, shape, **kw)
    sim = MDStyleSim(algo, Knobs(num_workers=W, num_servers=S, batch_size=B, img_shape=shape, **kw))
    sim.load(orc.net_g, orc.net_d)
    shared = []
    inner = sim.share_discriminators
    sim.share_discriminators = lambda: (shared.append(sim.t), inner())
    # the share alone, before any training: bit-exact
    before = sim.bank.rows().clone()
    sim.share_discriminators()
    orc.share()
    for c in range(W):
        ref = torch.cat([p.detach().reshape(-1) for p in orc.net_d[c].parameters()])
        assert torch.equal(sim.bank.rows()[c].cpu(), ref), (d_share, c)
    assert not torch.equal(sim.bank.rows(), before)
    shared.clear()
    for r in range(4):
        real, n_real, z_d, z_g = _round_inputs(W, S, B, d, 90 + r)
        l_ref = orc.round(real, n_real, z_d, z_g)
        l = sim.round(real.cuda(), n_real.cuda(), z_d.cuda(), z_g.cuda())
        assert (l.cpu() - l_ref).abs().max() < 1e-4, r
    assert shared == [0, 2]       # mdgan: t = 4, 2 (t % E == 0); acgan: rounds done 0, 2
    for c in range(W):
        ref = torch.cat([p.detach().reshape(-1) for p in orc.net_d[c].parameters()])
        assert_params_close(sim.bank.rows()[c], ref, steps=4, tag=(d_share, c), strict=False, bulk=1e-4)


def test_single_server_sharded_sim_on_one_rank_is_the_plain_sim(lib):
    """sim.MDSingleServerSim with world = 1 (no collective) is bit-identical to MDStyleSim with one server."""
    from cgl_gan_b200 import models
    from cgl_gan_b200.sim import Knobs, MDSingleServerSim, MDStyleSim
    torch.manual_seed(17)
    W, B, shape, d = 5, 100, (1, 28, 28), 784
    sizes = [300 + 41 * i for i in range(W)]
    k = Knobs(num_workers=W, num_servers=1, batch_size=B, img_shape=shape)
    a = MDStyleSim("mdgan", k, part_sizes=sizes)
    b = MDSingleServerSim("mdgan", k, part_sizes=sizes)
    g_mod = a.G.make_module()
    d_mods = [models.Discriminator(shape) for _ in range(W)]
    a.load([g_mod], d_mods)
    b.load([g_mod], d_mods)
    for r in range(2):
        real, n_real, z_d, z_g = _round_inputs(W, 1, B, d, 300 + r)
        la = a.round(real.cuda(), n_real.cuda(), z_d.cuda(), z_g.cuda())
        lb = b.round(real.cuda(), n_real.cuda(), z_d.cuda(), z_g.cuda())
        assert torch.equal(la, lb)
    assert torch.equal(a.bank.params, b.bank.params) and torch.equal(a.G.trunk.params, b.G.trunk.params)


def test_fl_local_epochs_are_full_unshuffled_passes(lib):
    """FLGAN-MNIST / FeGAN Worker.train: `epoch` full passes `for imgs in DataLoader(dataset, batch_size)` over the
    client's partition (FLGAN/MNIST/flgan.py:249-250); partitions of different sizes, short last batches."""
    from cgl_gan_b200.data import ResidentPartitions
    from cgl_gan_b200.sim import FLStyleSim, Knobs
    from oracle.rounds import OracleFL
    torch.manual_seed(23)
    C_, B, shape, d = 3, 100, (2,), 2
    data = torch.tanh(torch.randn(700, d))
    parts = [list(range(0, 250)), list(range(250, 330)), list(range(330, 700))]      # 3, 1 and 4 batches
    orc = OracleFL(C_, B, shape)
    orc.load_global()
    sim = FLStyleSim(Knobs(num_workers=C_, num_servers=1, batch_size=B, img_shape=shape))
    sim.load_global(orc.srv_g, orc.srv_d)
    rp = ResidentPartitions(data, parts, B, shuffle=False)
    gen = torch.Generator().manual_seed(4)
    zs = []

    def z_fn(n):
        z = (torch.randn(n, B, 100, generator=gen), torch.randn(n, B, 100, generator=gen))
        zs.append(z)
        return z[0].cuda(), z[1].cuda()

    done = sim.local_epochs(rp, epoch=2, z_fn=z_fn)
    assert done == 2 * (3 + 1 + 4)
    # the oracle: per client its own DataLoader loop, fed the noise the engine drew for that (minibatch, client)
    nb = [3, 1, 4]
    it = iter(zs)
    per_client = {c: [] for c in range(C_)}
    for _ in range(2):
        for j in range(max(nb)):
            z_d, z_g = next(it)
            active = [c for c in range(C_) if nb[c] > j]
            for a, c in enumerate(active):
                per_client[c].append((j, z_d[a], z_g[a]))
    from oracle import steps as st
    loss = st.make_loss(0)
    for c in range(C_):
        rows = torch.tensor(parts[c])
        for j, z_d, z_g in per_client[c]:
            imgs = data[rows[j * B:(j + 1) * B]]
            st.fl_local_minibatch(orc.net_d[c], orc.net_g[c], loss, orc.opti_g[c], orc.opti_d[c], imgs, z_d, z_g, B)
        ref = torch.cat([p.detach().reshape(-1) for p in orc.net_d[c].parameters()])
        assert_params_close(sim.bank.rows()[c], ref, steps=len(per_client[c]), tag=("D", c), strict=False, bulk=1e-4)
    assert sim.bank.step.tolist() == [6, 2, 8]
