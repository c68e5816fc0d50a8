"""Round-2 fixtures from the UNMODIFIED reference (/root/reference), generated in the build container:

  python tests/golden/make_golden2.py        -> tests/golden/steps2.json, steps2_arrays.npz

Pins the parts of the oracle that round 1 left unpinned ("oracle pinning is thin"):
  * Server.train of capgan.py:211-262, CAPGAN/MNIST/capgan.py:225-262, mixed-gan.py:238-292,
    MDGAN/MNIST/mdgan.py:180-207 and the multi-head MNIST branch of CGLGAN/MNIST/main.py:245-296 -- the bodies are
    typed here exactly as the reference states them (the scripts cannot be imported: matplotlib / fedlab / ignite,
    MNIST download and .cuda() at import time) and run on the reference's OWN model classes
    (imported by file path) with torch.optim / torch.nn on CPU;
  * one FL-GAN MNIST minibatch (FLGAN/MNIST/flgan.py:251-269) on FLGAN/MNIST/mnist_model.py;
  * Server.receive_parameter (CGLGAN/2DMG/main.py:171-179), lifted with `ast` and executed as is: the group mean;
  * the discriminator swap of MDGAN/MNIST/mdgan.py:158-164 (commented out as shipped, README.md:26): its three
    statements typed here, with the Server's own rd = Random(); rd.seed(rank + 100) (:122-123).
The GPU box has no /root/reference: tests only read the fixtures (tests/test_oracle_golden2.py).
"""
import ast
import json
import os
import sys
import types
from queue import Queue
from random import Random

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn, optim

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import REF, OUT, flat, lift, load, sha, summary  # noqa: E402

B = 100
IMS = (1, 28, 28)


def strided(t, n=4096):
    t = t.detach().reshape(-1)
    idx = torch.linspace(0, t.numel() - 1, min(n, t.numel())).long()
    return t[idx].numpy().copy()


def client_losses(net_ds, Xgs, kind):
    """Worker.train's tail on every client (capgan.py:343-347 / MDGAN/MNIST/mdgan.py:290-295), graph attached."""
    crit = nn.CrossEntropyLoss() if kind == "ce" else nn.BCELoss()
    out = []
    for net_d, Xg in zip(net_ds, Xgs):
        valid = torch.LongTensor(B).fill_(1) if kind == "ce" else torch.Tensor(B, 1).fill_(1)
        out.append(crit(net_d(Xg), valid))
    return out


def record(out, arrays, name, net_g, loss, F_max, Lambda, extra=None):
    stats = [v.reshape(-1) for k, v in net_g.state_dict().items() if "running" in k]
    out[name] = {"loss": [float(x) for x in loss.tolist()], "F_max": float(F_max), "Lambda": float(Lambda),
                 "g_params": summary(flat(net_g)), "stats": summary(torch.cat(stats)) if stats else None}
    if extra:
        out[name].update(extra)
    arrays[name + "_gparams"] = strided(flat(net_g))


def golden_server_updates():
    torch.set_num_threads(1)
    refmm = load("model/mnist_model.py", "ref_mm")
    refcgl = load("CGLGAN/MNIST/mnist_model.py", "ref_cgl_mnist")
    refmd = load("MDGAN/MNIST/mnist_model.py", "ref_md_mnist")
    reffl = load("FLGAN/MNIST/mnist_model.py", "ref_fl_mnist")
    out, arrays = {}, {}
    N = 2

    def inputs(seed):
        g = torch.Generator().manual_seed(seed)
        return torch.randn(B, 100, generator=g), torch.randn(B, 100, generator=g)

    # ---- capgan.py:211-262 ("exp weight") and CAPGAN/MNIST/capgan.py:225-262 ("mean weight") --------------------
    for name in ("cap_server", "cap_copy_server"):
        torch.manual_seed(20211212)
        net_g = refmm.Generator(IMS)
        net_ds = [refmm.Discriminator(IMS) for _ in range(N)]
        init = {"g_sha": sha(flat(net_g)), "d_sha": sha(flat(net_ds[0]))}
        opti = optim.Adam(net_g.parameters(), lr=0.0002, betas=(0.5, 0.999))
        Lambda = torch.tensor(0.7, requires_grad=True)            # capgan.py:159 starts at 0; 0.7 makes alpha non-uniform
        opti_L = optim.SGD([Lambda], lr=0.1)                      # capgan.py:160
        beta = torch.tensor([0.3, 0.7])
        z_d, z_g = inputs(21)
        with torch.no_grad():
            Xd = net_g(z_d)                                       # only its BatchNorm side effect survives
        Xg = net_g(z_g)
        opti.zero_grad()
        loss = torch.zeros(N)
        for i, g_loss in enumerate(client_losses(net_ds, [Xg.clone() for _ in range(N)], "ce")):
            loss[i] = g_loss.clone()
        opti_L.zero_grad()
        if name == "cap_server":
            alpha = F.softmax(Lambda.detach() * loss.detach(), dim=0)
            alpha = F.softmax(alpha * beta, dim=0)
            F_max = (alpha * loss).sum() - 0.001 * Lambda
        else:
            gamma = F.softmax(Lambda.detach() * loss.detach(), dim=0)
            s = F.softmax(beta * gamma, dim=0)
            F_max = (s * loss).sum() - 0.001 * Lambda
        F_max.backward()
        opti_L.step()
        opti.step()
        record(out, arrays, name, net_g, loss.detach(), F_max.item(), Lambda.item(), init)

    # ---- mixed-gan.py:238-292 on MixGenerator with weights_init (:68-77,181,348) -------------------------------
    ns = {"nn": nn}
    lift("mixed-gan.py", {"weights_init"}, ns)
    torch.manual_seed(20211212)
    net_g = refmm.MixGenerator(IMS, N)
    net_g.apply(ns["weights_init"])
    net_ds = [refmm.Discriminator(IMS) for _ in range(N)]
    for d_ in net_ds:
        d_.apply(ns["weights_init"])
    init = {"g_sha": sha(flat(net_g)), "d_sha": sha(flat(net_ds[0]))}
    opti = optim.Adam(net_g.parameters(), lr=0.0002, betas=(0.5, 0.999))
    Lambda = torch.tensor(0.7, requires_grad=True)
    opti_L = optim.SGD([Lambda], lr=0.1)
    beta = torch.tensor([0.3, 0.7])
    z_d, z_g = inputs(22)
    with torch.no_grad():
        Xd = torch.chunk(net_g(z_d), N, dim=0)
    Xg = torch.chunk(net_g(z_g), N, dim=0)
    opti.zero_grad()
    loss = torch.zeros(N)
    for i, g_loss in enumerate(client_losses(net_ds, [x.clone() for x in Xg], "ce")):
        loss[i] = g_loss.clone()
    losses = loss.sum()
    net_g.model.requires_grad_(False)
    losses.backward(retain_graph=True)
    net_g.model.requires_grad_(True)
    opti_L.zero_grad()
    alpha = F.softmax(beta * Lambda.detach() * loss.detach(), dim=0)
    F_max = (alpha * loss).sum() - 0.001 * Lambda
    net_g.paths.requires_grad_(False)
    F_max.backward()
    net_g.paths.requires_grad_(True)
    opti_L.step()
    opti.step()
    record(out, arrays, "mixed_server", net_g, loss.detach(), F_max.item(), Lambda.item(), init)

    # ---- MDGAN/MNIST/mdgan.py:180-207: losses = loss.mean() ------------------------------------------------------
    torch.manual_seed(20211212)
    net_g = refmd.Generator(IMS)
    net_ds = [refmd.Discriminator(IMS) for _ in range(3)]
    init = {"g_sha": sha(flat(net_g)), "d_sha": sha(flat(net_ds[0]))}
    opti = optim.Adam(net_g.parameters(), lr=0.0002, betas=(0.5, 0.999))
    z_d, z_g = inputs(23)
    with torch.no_grad():
        Xd = net_g(z_d)
    Xg = net_g(z_g)
    opti.zero_grad()
    loss = torch.zeros(3)
    for i, g_loss in enumerate(client_losses(net_ds, [Xg.clone() for _ in range(3)], "bce")):
        loss[i] = g_loss.clone()
    losses = loss.mean()
    losses.backward()
    opti.step()
    record(out, arrays, "mean_server", net_g, loss.detach(), losses.item(), 0.0, init)

    # ---- CGLGAN/MNIST/main.py:245-296, iid != 0 (multi-head) -----------------------------------------------------
    torch.manual_seed(20211212)
    net_g = refcgl.Generator(IMS, N)
    net_ds = [refcgl.Discriminator(IMS) for _ in range(N)]
    init = {"g_sha": sha(flat(net_g)), "d_sha": sha(flat(net_ds[0]))}
    opti = optim.Adam(net_g.parameters(), lr=0.0002, betas=(0.5, 0.999))
    Lambda = torch.tensor(0.5)
    beta = torch.tensor([0.3, 0.7])
    z_d, z_g = inputs(24)
    with torch.no_grad():
        Xd = torch.chunk(net_g(z_d), N, dim=0)
    Xg = torch.chunk(net_g(z_g), N, dim=0)
    opti.zero_grad()
    loss = torch.zeros(N)
    for i, g_loss in enumerate(client_losses(net_ds, [x.clone() for x in Xg], "bce")):
        loss[i] = g_loss.clone()
    losses = loss.sum()
    net_g.model.requires_grad_(False)
    losses.backward(retain_graph=True)
    net_g.model.requires_grad_(True)
    gamma = F.softmax(Lambda * loss, dim=0).detach()
    F_beta = (beta * loss).sum()
    F_gamma = (gamma * loss).sum()
    F_max = (F_beta + F_gamma) / 2
    net_g.paths.requires_grad_(False)
    F_max.backward()
    net_g.paths.requires_grad_(True)
    grad = (loss * loss * gamma).sum() - (loss * gamma * F_gamma).sum()
    Lambda = Lambda + 10 * grad
    opti.step()
    record(out, arrays, "cgl_mnist_server", net_g, loss.detach(), F_max.item(), Lambda.item(), init)

    # ---- FLGAN/MNIST/flgan.py:251-269: one minibatch of Worker.train ---------------------------------------------
    torch.manual_seed(20211212)
    net_g, net_d = reffl.Generator(IMS), reffl.Discriminator(IMS)
    init = {"g_sha": sha(flat(net_g)), "d_sha": sha(flat(net_d))}
    opti_g = optim.Adam(net_g.parameters(), lr=0.0002, betas=(0.5, 0.999))
    opti_d = optim.Adam(net_d.parameters(), lr=0.0002, betas=(0.5, 0.999))
    bce = nn.BCELoss()
    g = torch.Generator().manual_seed(25)
    imgs = torch.tanh(torch.randn(60, *IMS, generator=g))
    z_d, z_g = torch.randn(B, 100, generator=g), torch.randn(B, 100, generator=g)
    fake = torch.Tensor(B, 1).fill_(0)
    Xd = net_g(z_d)
    real_imgs = imgs.type(torch.FloatTensor)
    valid = torch.Tensor(imgs.shape[0], 1).fill_(1)
    opti_d.zero_grad()
    real_loss = bce(net_d(real_imgs), valid)
    fake_loss = bce(net_d(Xd), fake)
    D_loss = (real_loss + fake_loss)
    D_loss.backward()
    opti_d.step()
    valid = torch.Tensor(B, 1).fill_(1)
    opti_g.zero_grad()
    Xg = net_g(z_g)
    g_loss = bce(net_d(Xg), valid)
    g_loss.backward()
    opti_g.step()
    stats = torch.cat([v.reshape(-1) for k, v in net_g.state_dict().items() if "running" in k])
    out["fl_mnist"] = dict(init, d_loss=D_loss.item(), g_loss=g_loss.item(), d_params=summary(flat(net_d)),
                           g_params=summary(flat(net_g)), stats=summary(stats))
    arrays["fl_mnist_dparams"] = strided(flat(net_d))
    arrays["fl_mnist_gparams"] = strided(flat(net_g))
    torch.set_num_threads(os.cpu_count())
    return out, arrays


def lift_method(path, cls, name, ns):
    tree = ast.parse(open(os.path.join(REF, path)).read())
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for sub in node.body:
                if isinstance(sub, ast.FunctionDef) and sub.name == name:
                    exec(compile(ast.Module([sub], []), path, "exec"), ns)
                    return ns[name]
    raise KeyError((cls, name))


def golden_sharing():
    out, arrays = {}, {}
    # Server.receive_parameter, executed as written (CGLGAN/2DMG/main.py:171-179)
    recv = lift_method("CGLGAN/2DMG/main.py", "Server", "receive_parameter", {})
    g = torch.Generator().manual_seed(31)
    dicts = [{"model.0.weight": torch.randn(7, 5, generator=g), "model.0.bias": torch.randn(7, generator=g)} for _ in range(3)]
    self = types.SimpleNamespace(discriminator=Queue(), client_list=[4, 5, 6])
    for d_ in dicts:
        self.discriminator.put({k: v.clone() for k, v in d_.items()})
    p = recv(self)
    arrays["group_mean_in"] = torch.stack([torch.cat([d_["model.0.weight"].reshape(-1), d_["model.0.bias"]]) for d_ in dicts]).numpy()
    arrays["group_mean_out"] = torch.cat([p["model.0.weight"].reshape(-1), p["model.0.bias"]]).numpy()
    # the MD-GAN swap, MDGAN/MNIST/mdgan.py:158-164 with the Server's generator (:122-123)
    swaps = {}
    for rank in (0, 3):
        rd = Random()
        rd.seed(rank + 100)
        client_list = list(range(10))
        rounds = []
        for _ in range(3):
            p_ds = []
            for idx in client_list:
                p_ds.append(idx)                 # stands for queen_d.get(): worker idx's state dict
            rd.shuffle(p_ds)
            rounds.append([p_ds[idx] for idx in client_list])   # workers[idx].para_d.put(p_ds[idx])
        swaps[str(rank)] = rounds
    out["swap"] = swaps
    return out, arrays


def golden_conv_lsgan():
    """model/lsgan.py Generator / Discriminator (no call site in the scripts): initial weights by hash, one training-mode
    forward of each (BatchNorm2d batch statistics; Dropout2d noise from the seeded global RNG), one eval-mode forward."""
    torch.set_num_threads(1)
    ref = load("model/lsgan.py", "ref_lsgan")
    torch.manual_seed(20211212)
    net_g, net_d = ref.Generator(None), ref.Discriminator(None)
    out = {"g_sha": sha(flat(net_g)), "d_sha": sha(flat(net_d))}
    g = torch.Generator().manual_seed(41)
    z = torch.randn(4, 100, generator=g)
    img = net_g(z)
    torch.manual_seed(5)
    val = net_d(img.detach())
    out["img_train"] = summary(img)
    out["val_train"] = [float(v) for v in val.reshape(-1).tolist()]
    out["g_stats"] = summary(torch.cat([b.reshape(-1) for n, b in net_g.named_buffers() if "running" in n]))
    out["d_stats"] = summary(torch.cat([b.reshape(-1) for n, b in net_d.named_buffers() if "running" in n]))
    net_g.eval(); net_d.eval()
    with torch.no_grad():
        img_e = net_g(z)
        val_e = net_d(img_e)
    out["img_eval"] = summary(img_e)
    out["val_eval"] = [float(v) for v in val_e.reshape(-1).tolist()]
    torch.set_num_threads(os.cpu_count())
    return {"conv_lsgan": out}


if __name__ == "__main__":
    steps, arrays = golden_server_updates()
    s2, a2 = golden_sharing()
    steps.update(s2)
    arrays.update(a2)
    steps.update(golden_conv_lsgan())
    json.dump(steps, open(os.path.join(OUT, "steps2.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(OUT, "steps2_arrays.npz"), **arrays)
    print("wrote steps2.json, steps2_arrays.npz:", sorted(steps))
