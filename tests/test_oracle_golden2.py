"""Pins the rest of the oracle against reference-run fixtures (tests/golden/make_golden2.py): the Server.train
variants of capgan.py / CAPGAN/MNIST/capgan.py / mixed-gan.py / MDGAN/MNIST/mdgan.py / CGLGAN/MNIST/main.py on the
reference's own MNIST classes, one FL-GAN MNIST minibatch, receive_parameter's group mean and the MD-GAN swap. CPU only."""
import hashlib
import json
import os
import random

import numpy as np
import pytest
import torch

from oracle import models as om
from oracle import steps as st

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STEPS = json.load(open(os.path.join(GOLD, "steps2.json")))
ARR = np.load(os.path.join(GOLD, "steps2_arrays.npz"))
B, IMS = 100, (1, 28, 28)


def _flat(net):
    return torch.cat([p.detach().reshape(-1) for p in net.parameters()])


def _sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes()).hexdigest()[:16]


def _strided(t, n=4096):
    t = t.detach().reshape(-1)
    idx = torch.linspace(0, t.numel() - 1, min(n, t.numel())).long()
    return t[idx]


def _check_summary(t, s, tol=1e-6):
    t = t.detach().double().reshape(-1)
    assert t.numel() == s["numel"]
    idx = torch.linspace(0, t.numel() - 1, 64).long()
    scale = max(abs(x) for x in s["samples"]) + 1e-30
    assert max(abs(a - b) for a, b in zip(t[idx].tolist(), s["samples"])) <= tol * scale
    assert abs(t.abs().sum().item() - s["abs_sum"]) <= 1e-6 * s["abs_sum"] + 1e-9


@pytest.fixture(autouse=True)
def one_thread():
    """The fixtures were produced single-threaded: the same summation order, so the comparison can be tight."""
    keep = torch.get_num_threads()
    torch.set_num_threads(1)
    yield
    torch.set_num_threads(keep)


def _inputs(seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 100, generator=g), torch.randn(B, 100, generator=g)


def _losses(net_ds, Xg, kind):
    crit = st.make_loss(kind)
    loss = torch.zeros(len(net_ds))
    for i, d in enumerate(net_ds):
        loss[i] = st.worker_g_loss(d, crit, kind, Xg[i].clone(), B).clone()
    return loss


def _finish(name, net_g, loss, F_max, Lambda):
    gold = STEPS[name]
    assert np.allclose(loss.tolist(), gold["loss"], atol=1e-6), (loss.tolist(), gold["loss"])
    assert abs(float(F_max) - gold["F_max"]) < 1e-6
    assert abs(float(Lambda) - gold["Lambda"]) < 1e-6
    _check_summary(_flat(net_g), gold["g_params"])
    stats = torch.cat([v.reshape(-1) for k, v in net_g.state_dict().items() if "running" in k])
    _check_summary(stats, gold["stats"])
    ref = torch.from_numpy(ARR[name + "_gparams"])
    got = _strided(_flat(net_g))
    assert (got - ref).abs().max().item() <= 1e-6 * ref.abs().max().item(), name


@pytest.mark.parametrize("name,update", [("cap_server", st.server_update_capgan),
                                         ("cap_copy_server", st.server_update_capgan_copy)])
def test_capgan_server_updates_match_reference(name, update):
    gold = STEPS[name]
    torch.manual_seed(20211212)
    net_g = om.GeneratorMNIST(IMS)
    net_ds = [om.DiscriminatorMNIST2(IMS) for _ in range(2)]
    assert _sha(_flat(net_g)) == gold["g_sha"] and _sha(_flat(net_ds[0])) == gold["d_sha"]
    opti = st.make_adam(net_g.parameters())
    Lambda = torch.tensor(0.7, requires_grad=True)
    opti_L = torch.optim.SGD([Lambda], lr=0.1)
    z_d, z_g = _inputs(21)
    Xd, Xg = st.server_generate(net_g, z_d, z_g, 2, multi_head=False)
    loss = _losses(net_ds, Xg, st.LOSS_CE)
    F_max = update(net_g, opti, loss, torch.tensor([0.3, 0.7]), Lambda, opti_L)
    _finish(name, net_g, loss.detach(), F_max, Lambda.item())


def test_mixed_server_update_matches_reference():
    gold = STEPS["mixed_server"]
    torch.manual_seed(20211212)
    net_g = om.MixGeneratorMNIST(IMS, 2)
    net_g.apply(om.weights_init)
    net_ds = [om.DiscriminatorMNIST2(IMS) for _ in range(2)]
    for d in net_ds:
        d.apply(om.weights_init)
    assert _sha(_flat(net_g)) == gold["g_sha"] and _sha(_flat(net_ds[0])) == gold["d_sha"]
    opti = st.make_adam(net_g.parameters())
    Lambda = torch.tensor(0.7, requires_grad=True)
    opti_L = torch.optim.SGD([Lambda], lr=0.1)
    z_d, z_g = _inputs(22)
    Xd, Xg = st.server_generate(net_g, z_d, z_g, 2, multi_head=True)
    loss = _losses(net_ds, Xg, st.LOSS_CE)
    F_max = st.server_update_mixed(net_g, opti, loss, torch.tensor([0.3, 0.7]), Lambda, opti_L)
    _finish("mixed_server", net_g, loss.detach(), F_max, Lambda.item())


def test_mean_server_update_matches_reference():
    gold = STEPS["mean_server"]
    torch.manual_seed(20211212)
    net_g = om.GeneratorMNIST(IMS)
    net_ds = [om.DiscriminatorMNIST1(IMS) for _ in range(3)]
    assert _sha(_flat(net_g)) == gold["g_sha"] and _sha(_flat(net_ds[0])) == gold["d_sha"]
    opti = st.make_adam(net_g.parameters())
    z_d, z_g = _inputs(23)
    Xd, Xg = st.server_generate(net_g, z_d, z_g, 3, multi_head=False)
    loss = _losses(net_ds, Xg, st.LOSS_BCE)
    F_max = st.server_update_mean(net_g, opti, loss)
    _finish("mean_server", net_g, loss.detach(), F_max, 0.0)


def test_cglgan_mnist_multi_head_update_matches_reference():
    gold = STEPS["cgl_mnist_server"]
    torch.manual_seed(20211212)
    net_g = om.MixGeneratorMNIST(IMS, 2)
    net_ds = [om.DiscriminatorMNIST1(IMS) for _ in range(2)]
    assert _sha(_flat(net_g)) == gold["g_sha"] and _sha(_flat(net_ds[0])) == gold["d_sha"]
    opti = st.make_adam(net_g.parameters())
    z_d, z_g = _inputs(24)
    Xd, Xg = st.server_generate(net_g, z_d, z_g, 2, multi_head=True)
    loss = _losses(net_ds, Xg, st.LOSS_BCE)
    lam, F_max = st.server_update_cglgan(net_g, opti, loss, torch.tensor([0.3, 0.7]), torch.tensor(0.5), True)
    _finish("cgl_mnist_server", net_g, loss.detach(), F_max, lam.item())


def test_fl_mnist_minibatch_matches_reference():
    gold = STEPS["fl_mnist"]
    torch.manual_seed(20211212)
    net_g, net_d = om.GeneratorMNIST(IMS), om.DiscriminatorMNIST1(IMS)
    assert _sha(_flat(net_g)) == gold["g_sha"] and _sha(_flat(net_d)) == gold["d_sha"]
    opti_g, opti_d = st.make_adam(net_g.parameters()), st.make_adam(net_d.parameters())
    g = torch.Generator().manual_seed(25)
    imgs = torch.tanh(torch.randn(60, *IMS, generator=g))
    z_d, z_g = torch.randn(B, 100, generator=g), torch.randn(B, 100, generator=g)
    dl, gl = st.fl_local_minibatch(net_d, net_g, st.make_loss(st.LOSS_BCE), opti_g, opti_d, imgs, z_d, z_g, B)
    assert abs(dl.item() - gold["d_loss"]) < 1e-6 and abs(gl.item() - gold["g_loss"]) < 1e-6
    _check_summary(_flat(net_d), gold["d_params"])
    _check_summary(_flat(net_g), gold["g_params"])
    stats = torch.cat([v.reshape(-1) for k, v in net_g.state_dict().items() if "running" in k])
    _check_summary(stats, gold["stats"])
    for key, net in (("fl_mnist_dparams", net_d), ("fl_mnist_gparams", net_g)):
        ref = torch.from_numpy(ARR[key])
        assert (_strided(_flat(net)) - ref).abs().max().item() <= 1e-6 * ref.abs().max().item()


def test_group_mean_is_receive_parameter():
    """oracle.steps.group_mean == Server.receive_parameter executed as written, bit for bit."""
    rows = torch.from_numpy(ARR["group_mean_in"])
    out = st.group_mean([{"w": r.clone()} for r in rows])["w"]
    assert torch.equal(out, torch.from_numpy(ARR["group_mean_out"]))


def test_mdgan_swap_matches_reference_statements():
    for rank, rounds in STEPS["swap"].items():
        rd = random.Random(int(rank) + 100)
        for expect in rounds:
            assert st.mdgan_swap(list(range(10)), rd) == expect


def test_conv_lsgan_classes_match_reference():
    """oracle.models.ConvGenerator / ConvDiscriminator == model/lsgan.py's classes: same initial weights for a seed, the same
    training-mode forward (BatchNorm2d batch statistics, Dropout2d noise from the seeded global RNG) and eval-mode forward."""
    gold = STEPS["conv_lsgan"]
    torch.manual_seed(20211212)
    net_g, net_d = om.ConvGenerator(None), om.ConvDiscriminator(None)
    assert _sha(_flat(net_g)) == gold["g_sha"] and _sha(_flat(net_d)) == gold["d_sha"]
    g = torch.Generator().manual_seed(41)
    z = torch.randn(4, 100, generator=g)
    img = net_g(z)
    torch.manual_seed(5)
    val = net_d(img.detach())
    _check_summary(img, gold["img_train"])
    assert np.allclose(val.reshape(-1).tolist(), gold["val_train"], atol=1e-6)
    _check_summary(torch.cat([b.reshape(-1) for n, b in net_g.named_buffers() if "running" in n]), gold["g_stats"])
    _check_summary(torch.cat([b.reshape(-1) for n, b in net_d.named_buffers() if "running" in n]), gold["d_stats"])
    # the injected-mask path is the same forward: the masks one training-mode forward draws, fed back in
    torch.manual_seed(5)
    masks = om.draw_dropout2d_masks(4)
    net_d2 = om.ConvDiscriminator(None)
    net_d2.load_state_dict({k: v for k, v in net_d.state_dict().items()})
    torch.manual_seed(5)
    a = net_d2(img.detach())
    net_d3 = om.ConvDiscriminator(None)
    net_d3.load_state_dict({k: v for k, v in net_d.state_dict().items()})
    b = net_d3(img.detach(), masks)
    assert torch.equal(a, b)
    net_g.eval(); net_d.eval()
    with torch.no_grad():
        img_e = net_g(z)
        val_e = net_d(img_e)
    _check_summary(img_e, gold["img_eval"])
    assert np.allclose(val_e.reshape(-1).tolist(), gold["val_eval"], atol=1e-6)
