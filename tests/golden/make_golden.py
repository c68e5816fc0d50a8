"""Generates tests/golden/*.npz|json from the UNMODIFIED reference (/root/reference), in the build container.

  python tests/golden/make_golden.py

What is taken from the reference itself:
  * its model classes (imported by file path) -> initial parameters for a seed, forward outputs;
  * its step bodies -- Worker.train's D loop and G loss (CGLGAN/2DMG/main.py:357-372, capgan.py:329-346),
    FL Worker.train's minibatch (FLGAN/2DMG/flgan.py:239-256), Server.train's CGLGAN update
    (CGLGAN/2DMG/main.py:245-276) -- typed here exactly as the reference states them, run on the reference's
    classes with torch.optim.Adam / nn.BCELoss / nn.CrossEntropyLoss on CPU;
  * its pure functions allocate_dataset / del_tensor_ele / init_groups, lifted verbatim with `ast` and
    executed against the reference's own global names.
The driver scripts themselves cannot be imported (matplotlib / fedlab / ignite missing, MNIST download,
hard-coded .cuda()). The GPU box has no /root/reference: tests only read the fixtures.
"""
import ast
import copy
import hashlib
import importlib.util
import json
import os
import sys
from queue import Queue
from random import Random

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn, optim

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, path))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def lift(path, names, ns):
    tree = ast.parse(open(os.path.join(REF, path)).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes()).hexdigest()[:16]


def flat(net):
    return torch.cat([p.detach().reshape(-1) for p in net.parameters()])


def summary(t):
    t = t.detach().double().reshape(-1)
    idx = torch.linspace(0, t.numel() - 1, 64).long()
    return {"sum": t.sum().item(), "abs_sum": t.abs().sum().item(), "sq_sum": (t * t).sum().item(),
            "samples": t[idx].tolist(), "numel": t.numel()}


# ---------------------------------------------------------------------------------------------------
def golden_steps():
    torch.set_num_threads(1)   # fixed summation order for the fixture
    ref2d = load("CGLGAN/2DMG/model.py", "ref_2d")
    refmn = load("CGLGAN/MNIST/mnist_model.py", "ref_cgl_mnist")
    refmm = load("model/mnist_model.py", "ref_mm")
    refmd2 = load("MDGAN/2DMG/model.py", "ref_md2d")
    out = {}
    B = 100
    cases = [("d2d_bce", lambda: ref2d.Discriminator(), 2, "bce", 1.0),
             ("dmnist1_bce", lambda: refmn.Discriminator((1, 28, 28)), 784, "bce", 1.0),
             ("dmnist2_ce", lambda: refmm.Discriminator((1, 28, 28)), 784, "ce", 0.5)]
    arrays = {}
    for name, mk, d, kind, scale in cases:
        torch.manual_seed(20211212)
        net_d = mk()
        init_sha = sha(flat(net_d))
        g = torch.Generator().manual_seed(77)
        imgs = torch.tanh(torch.randn(41, d, generator=g))     # a short last batch
        X = torch.tanh(torch.randn(B, d, generator=g) * 0.5)
        Xg = torch.tanh(torch.randn(B, d, generator=g) * 0.5).requires_grad_(True)
        loss = nn.BCELoss() if kind == "bce" else nn.CrossEntropyLoss()
        opti_d = optim.Adam(net_d.parameters(), lr=0.0002, betas=(0.5, 0.999))
        rec = {"init_sha": init_sha, "d_loss": [], "g_loss": []}
        for it in range(2):
            # --- CGLGAN/2DMG/main.py:357-366 / capgan.py:329-341 ---
            if kind == "bce":
                valid = torch.Tensor(imgs.shape[0], 1).fill_(1)
                fake = torch.Tensor(B, 1).fill_(0)
            else:
                valid = torch.LongTensor(imgs.shape[0]).fill_(1)
                fake = torch.LongTensor(B).fill_(0)
            real_imgs = imgs.type(torch.FloatTensor)
            opti_d.zero_grad()
            real_loss = loss(net_d(real_imgs), valid)
            fake_loss = loss(net_d(X), fake)
            D_loss = (real_loss + fake_loss) if kind == "bce" else (real_loss + fake_loss) * 0.5
            D_loss.backward()
            opti_d.step()
            # --- CGLGAN/2DMG/main.py:368-372 ---
            valid = torch.Tensor(B, 1).fill_(1) if kind == "bce" else torch.LongTensor(B).fill_(1)
            Xg.grad = None
            G_loss = loss(net_d(Xg), valid)
            G_loss.backward()
            rec["d_loss"].append(D_loss.item())
            rec["g_loss"].append(G_loss.item())
        rec["params"] = summary(flat(net_d))
        rec["dxg"] = summary(Xg.grad)
        if d == 2:
            arrays[name + "_params"] = flat(net_d).numpy()
            arrays[name + "_dxg"] = Xg.grad.numpy()
        out[name] = rec

    # generator forward + CGLGAN server update on the reference's multi-head 2-D generator
    torch.manual_seed(20211212)
    net_g = ref2d.Generator((2,), 2)
    net_ds = [ref2d.Discriminator() for _ in range(2)]
    opti = optim.Adam(net_g.parameters(), lr=0.0002, betas=(0.5, 0.999))
    g = torch.Generator().manual_seed(5)
    z = torch.randn(B, 100, generator=g)
    Xg = torch.chunk(net_g(z), 2, dim=0)
    bce = nn.BCELoss()
    beta = torch.tensor([0.3, 0.7])
    Lambda = torch.tensor(0.5)
    opti.zero_grad()
    loss = torch.zeros(2)
    for i in range(2):
        loss[i] = bce(net_ds[i](Xg[i]), torch.ones(B, 1)).clone()
    # --- CGLGAN/2DMG/main.py:254-276 (iid != 0) ---
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
    out["cgl_server_2d"] = {"loss": loss.tolist(), "F_max": F_max.item(), "Lambda": Lambda.item(),
                            "g_params": summary(flat(net_g))}
    arrays["cgl_server_2d_gparams"] = flat(net_g).numpy()

    # MNIST generators: forward in train mode (BN eps 0.8 batch statistics) twice, running stats
    for name, mk in [("g_mnist", lambda: refmm.Generator((1, 28, 28))),
                     ("mixg_mnist", lambda: refmm.MixGenerator((1, 28, 28), 2))]:
        torch.manual_seed(20211212)
        net = mk()
        g = torch.Generator().manual_seed(9)
        z = torch.randn(16, 100, generator=g)
        net(z)
        y = net(z)
        stats = torch.cat([v.reshape(-1) for k, v in net.state_dict().items() if "running" in k])
        out[name] = {"init_sha": sha(flat(net)), "y": summary(y), "stats": summary(stats)}

    # FL minibatch on the reference 2-D classes (FLGAN/2DMG/flgan.py:239-256)
    torch.manual_seed(20211212)
    net_g, net_d = refmd2.Generator((2,)), refmd2.Discriminator()
    opti_g = optim.Adam(net_g.parameters(), lr=0.0002, betas=(0.5, 0.999))
    opti_d = optim.Adam(net_d.parameters(), lr=0.0002, betas=(0.5, 0.999))
    g = torch.Generator().manual_seed(3)
    imgs = torch.tanh(torch.randn(60, 2, generator=g))
    fake = torch.Tensor(B, 1).fill_(0)
    rec = {"d_loss": [], "g_loss": []}
    for it in range(2):
        valid = torch.Tensor(imgs.shape[0], 1).fill_(1)
        z = torch.randn(B, 100, generator=g)
        Xd = net_g(z)
        real_imgs = imgs.type(torch.FloatTensor)
        opti_d.zero_grad()
        real_loss = bce(net_d(real_imgs), valid)
        fake_loss = bce(net_d(Xd), fake)
        D_loss = (real_loss + fake_loss)
        D_loss.backward()
        opti_d.step()
        valid = torch.Tensor(B, 1).fill_(1)
        opti_g.zero_grad()
        z = torch.randn(B, 100, generator=g)
        Xgg = net_g(z)
        g_loss = bce(net_d(Xgg), valid)
        g_loss.backward()
        opti_g.step()
        rec["d_loss"].append(D_loss.item()); rec["g_loss"].append(g_loss.item())
    rec["d_params"] = summary(flat(net_d)); rec["g_params"] = summary(flat(net_g))
    arrays["fl2d_d_params"] = flat(net_d).numpy(); arrays["fl2d_g_params"] = flat(net_g).numpy()
    out["fl2d"] = rec
    torch.set_num_threads(os.cpu_count())
    return out, arrays


# ---------------------------------------------------------------------------------------------------
class _FakeTV:
    """Minimal stand-in for a torchvision dataset: .data (here: the sample's original index) and .targets."""

    def __init__(self, data, targets):
        self.data, self.targets = data, targets

    def __len__(self):
        return len(self.targets)


def golden_partitions():
    out = {}
    rs = np.random.RandomState(1)
    n, num_class = 3000, 10
    labels = torch.from_numpy(rs.randint(0, num_class, size=n)).long()
    data = torch.arange(n).float().unsqueeze(1)       # each sample carries its own index

    class DS(torch.utils.data.Dataset):
        def __len__(self): return n
        def __getitem__(self, i): return data[i], labels[i]

    for style, path, nw in [("cgl", "CGLGAN/2DMG/main.py", 10), ("fl", "FLGAN/MNIST/flgan.py", 10),
                            ("fl2d", "FLGAN/2DMG/flgan.py", 10)]:
        rd = Random(); rd.seed(20211212)
        ns = {"torch": torch, "np": np, "copy": copy, "rd": rd, "num_workers": nw, "num_class": num_class,
              "num_sample": 100, "datasets": [], "test_set": [], "ims": 0}
        lift(path, {"allocate_dataset", "del_tensor_ele"}, ns)
        rec = {}
        for iid in (0, 1, 2):             # the reference loops iid with ONE continuing rd stream
            ns["datasets"].clear()
            ns["allocate_dataset"](DS(), iid)
            rec[str(iid)] = {"parts": [[int(v) for v in d_.reshape(-1).tolist()] for d_ in ns["datasets"]],
                             "test": [int(v) for v in ns["test_set"].reshape(-1).tolist()][:100] if style != "fl2d" else None}
        out[style] = rec
    # torchvision form (capgan.py:358-424)
    rd = Random(); rd.seed(20211212)
    ns = {"torch": torch, "np": np, "copy": copy, "rd": rd, "num_workers": 10, "num_class": num_class,
          "num_sample": 100, "datasets": [], "test_set": []}
    lift("capgan.py", {"allocate_dataset", "del_tensor_ele"}, ns)
    rec = {}
    for iid in (0, 1, 2):
        ns["datasets"].clear()
        ds = _FakeTV(torch.arange(n), labels.clone())
        ns["allocate_dataset"](ds, iid)
        rec[str(iid)] = {"parts": [[int(v) for v in d_.data.tolist()] for d_ in ns["datasets"]],
                         "test": [int(v) for v in ns["test_set"].tolist()]}
    out["cap"] = rec
    out["labels"] = labels.tolist()

    # the survey's known-answer sizes for the repo-default CGLGAN 2DMG run
    sys.path.insert(0, os.path.join(REF, "CGLGAN/2DMG"))
    np.random.seed(20211212); torch.manual_seed(20211212)
    gm = load("CGLGAN/2DMG/data.py", "ref_data").gmm(10, 10000)
    counts = [int((gm.targets == c).sum()) for c in range(10)]
    rd = Random(); rd.seed(20211212)
    ns = {"torch": torch, "np": np, "copy": copy, "rd": rd, "num_workers": 10, "num_class": 10,
          "num_sample": 10000, "datasets": [], "test_set": [], "ims": 0}
    lift("CGLGAN/2DMG/main.py", {"allocate_dataset", "del_tensor_ele"}, ns)
    sizes = {}
    for iid in (0, 1, 2):
        ns["datasets"].clear()
        ns["allocate_dataset"](gm, iid)
        sizes[str(iid)] = [len(d_) for d_ in ns["datasets"]]
        for i in range(10):                       # main.py:461 -- the save_image sample between workers
            rd.sample(range(len(ns["datasets"][i])), 100)
    out["gmm_default"] = {"class_counts": counts, "sizes": sizes, "labels_sha": sha(gm.targets)}
    np.save(os.path.join(OUT, "gmm_labels.npy"), gm.targets.numpy().astype(np.int8))

    # init_groups (fegan.py:383-452)
    for frac, tag in [(0.2, "frac02"), (1, "frac1")]:
        ns = {"np": np, "Queue": Queue, "frac_workers": frac}
        lift("fegan.py", {"init_groups"}, ns)
        xs = [np.eye(10, dtype=np.int64)[i] * 100 for i in range(10)]
        g1 = ns["init_groups"](10, xs)[:40]
        ns = {"np": np, "Queue": Queue, "frac_workers": frac}
        lift("fegan.py", {"init_groups"}, ns)
        rs = np.random.RandomState(4)
        xs2 = [rs.randint(0, 3, size=10) * rs.randint(1, 50, size=10) for _ in range(10)]
        for x in xs2:
            if x.sum() == 0:
                x[0] = 5
        g2 = ns["init_groups"](10, xs2)[:40]
        out["init_groups_" + tag] = {"onehot": [[int(v) for v in g] for g in g1],
                                     "random": [[int(v) for v in g] for g in g2],
                                     "random_freq": [[int(v) for v in x] for x in xs2]}
    return out


if __name__ == "__main__":
    steps, arrays = golden_steps()
    json.dump(steps, open(os.path.join(OUT, "steps.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(OUT, "steps_arrays.npz"), **arrays)
    parts = golden_partitions()
    json.dump(parts, open(os.path.join(OUT, "partitions.json"), "w"))
    print("wrote", os.listdir(OUT))
