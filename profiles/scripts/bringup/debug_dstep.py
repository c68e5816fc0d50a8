"""Bring-up diagnostic (not a pytest file): one discriminator step in FFMA mode vs tcgen05 mode from the same
state, element-wise, per parameter tensor; and both against the CPU oracle."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import make_batches, make_ds, osteps
from cgl_gan_b200 import abi
from cgl_gan_b200.engine import ClientBank

arch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
G, B = 5, 100
kind = {0: 0, 1: 0, 2: 1, 3: 2}[arch]
nets = make_ds(arch, G, seed=100 + arch)
real, fake, xg = make_batches(arch, G, B, seed=arch)
n_real = torch.tensor([B, 41, 1, B, 77])
for g in range(G):
    real[g, n_real[g]:] = 0
res = {}
for mode in (1, 2):
    abi.check(abi.lib.cgl_set_gemm_mode(mode))
    bank = ClientBank(arch, G, B, loss_kind=kind)
    bank.load_modules(nets)
    bank.d_step(real.cuda(), fake.cuda(), n_real=n_real)
    l, dx = bank.g_loss_raw(xg.cuda())
    torch.cuda.synchronize()
    res[mode] = (bank.rows().cpu().clone(), dx.cpu().clone(), bank.lay)
loss = osteps.make_loss(kind)
ref_rows, ref_dx = [], []
for g in range(G):
    opt = osteps.make_adam(nets[g].parameters())
    osteps.worker_d_step(nets[g], opt, loss, kind, real[g, :n_real[g]], fake[g], B)
    x = xg[g].clone().requires_grad_(True)
    osteps.worker_g_loss(nets[g], loss, kind, x, B).backward()
    ref_rows.append(torch.cat([p.detach().reshape(-1) for p in nets[g].parameters()]))
    ref_dx.append(x.grad)
ref_rows, ref_dx = torch.stack(ref_rows), torch.stack(ref_dx)
lay = res[1][2]
segs = []
for l in range(lay.n_layers):
    segs.append((f"W{l}", lay.w_off[l], lay.dims[l] * lay.dims[l + 1]))
    segs.append((f"b{l}", lay.b_off[l], lay.dims[l + 1]))
for name, a, b in (("ffma vs ref", res[1][0], ref_rows), ("tc vs ref", res[2][0], ref_rows), ("tc vs ffma", res[2][0], res[1][0])):
    print(name)
    for sname, off, n in segs:
        d = (a[:, off:off + n] - b[:, off:off + n]).abs()
        print(f"   {sname}: max {d.max():.3e}  frac>1e-7 {(d > 1e-7).float().mean():.3e}  frac>1e-4 {(d > 1e-4).float().mean():.3e}")
for name, a in (("ffma", res[1][1]), ("tc", res[2][1])):
    for g in range(G):
        e = (a[g] - ref_dx[g]).abs().amax(dim=1) / ref_dx[g].abs().amax()
        print(name, g, "dxg rows > 1e-5:", (e > 1e-5).float().mean().item(), " > 1e-4:", (e > 1e-4).float().mean().item(), "max", e.max().item(),
              "rowmax ref", ref_dx[g].abs().amax(dim=1).median().item(), ref_dx[g].abs().amax().item())
