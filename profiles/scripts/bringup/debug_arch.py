"""Manual GPU debugging aid: stage-by-stage comparison of one D step + G loss against the oracle."""
import sys
import torch
from helpers import make_batches, make_ds, rel_err, rel_l2, max_abs, osteps

arch = int(sys.argv[1]) if len(sys.argv) > 1 else 3
kind = {0: 0, 1: 0, 2: 1, 3: 2}[arch]
scale = 0.5 if arch == 2 else 1.0
from cgl_gan_b200.engine import ClientBank
G, B = 5, 100
nets = make_ds(arch, G, seed=100 + arch)
bank = ClientBank(arch, G, B, loss_kind=kind, d_loss_scale=scale)
bank.load_modules(nets)
loss = osteps.make_loss(kind)
optis = [osteps.make_adam(n.parameters()) for n in nets]
real, fake, xg = make_batches(arch, G, B, seed=arch)
n_real = torch.tensor([B, 41, 1, B, 77])
real_pad = real.clone()
for g in range(G):
    real_pad[g, n_real[g]:] = 0
# G loss BEFORE the step (isolates g_loss from d_step)
l0, dx0 = bank.g_loss_raw(xg.cuda())
for g in range(G):
    x = xg[g].clone().requires_grad_(True)
    l = osteps.worker_g_loss(nets[g], loss, kind, x, B); l.backward()
    print(f"pre-step g={g} gloss {l0[g].item():.7f} vs {l.item():.7f}  dxg rel_err {rel_err(dx0[g], x.grad):.3e}")
d_gpu = bank.d_step(real_pad.cuda(), fake.cuda(), n_real=n_real)
l1, dx1 = bank.g_loss_raw(xg.cuda())
for g in range(G):
    d_ref = osteps.worker_d_step(nets[g], optis[g], loss, kind, real[g, :n_real[g]], fake[g], B, scale)
    x = xg[g].clone().requires_grad_(True)
    l = osteps.worker_g_loss(nets[g], loss, kind, x, B); l.backward()
    print(f"g={g} dloss {d_gpu[g].item():.7f} vs {d_ref.item():.7f} | gloss {l1[g].item():.7f} vs {l.item():.7f} dxg rel_err {rel_err(dx1[g], x.grad):.3e}")
    off = 0
    for name, p in nets[g].named_parameters():
        n = p.numel()
        a = bank.rows()[g, off:off + n]
        print(f"    {name:16s} rel_l2 {rel_l2(a, p.reshape(-1)):.3e} max_abs {max_abs(a, p.reshape(-1)):.3e}")
        off += n

# ---- locate the worst element of client 0's first layer and recompute its gradient in fp64 ----
import copy
g = 0
nets0 = make_ds(arch, G, seed=100 + arch)
net64 = copy.deepcopy(nets0[g]).double()
loss64 = osteps.make_loss(kind)
r64 = real[g, :n_real[g]].double(); f64 = fake[g].double()
tv = osteps._targets(kind, r64.shape[0], 1); tf = osteps._targets(kind, B, 0)
if kind != 1:
    tv, tf = tv.double(), tf.double()
L = (loss64(net64(r64), tv) + loss64(net64(f64), tf)) * scale
L.backward()
g64 = net64.model[0].weight.grad.reshape(-1)
p_init = nets0[g].model[0].weight.detach().reshape(-1)
p_cpu = nets[g].model[0].weight.detach().reshape(-1)
n0 = p_cpu.numel()
p_gpu = bank.rows()[g, :n0].cpu()
diff = (p_gpu - p_cpu).abs()
top = torch.topk(diff, 8).indices
for i in top.tolist():
    print(f"idx {i} (o={i // 784}, i={i % 784}) init {p_init[i]:.8f} cpu {p_cpu[i]:.8f} gpu {p_gpu[i]:.8f} "
          f"dcpu {(p_cpu[i]-p_init[i]).item():+.3e} dgpu {(p_gpu[i]-p_init[i]).item():+.3e} g64 {g64[i].item():+.3e}")
print("count |diff|>1e-5:", int((diff > 1e-5).sum()), " |g64|<1e-7:", int((g64.abs() < 1e-7).sum()))
