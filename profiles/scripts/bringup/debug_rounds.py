"""Bring-up diagnostic (not a pytest file): one CASE of test_gpu_rounds run with the FFMA kernel and with the
automatic kernel choice, compared with each other and with the oracle after every round."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import quantile_err, max_abs
from oracle.rounds import OracleMD
from test_gpu_rounds import CASES, _inputs
from cgl_gan_b200 import abi
from cgl_gan_b200.sim import Knobs, MDStyleSim

case = CASES[int(sys.argv[1]) if len(sys.argv) > 1 else 0]
algo, shape, W, S, iid, segema, epoch, rounds = case
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else rounds
B, d = 100, 1
for s in shape:
    d *= s
sizes = [1000 + 137 * i for i in range(W)]
torch.manual_seed(20211212)
orc = OracleMD(algo, W, S, B, shape, iid=iid, part_sizes=sizes, segema=segema, weights_init=(algo == "mixed"))
k = Knobs(num_workers=W, num_servers=S, batch_size=B, epoch=epoch, segema=segema, iid=iid, img_shape=shape)
sims = {}
for mode in (1, 0):
    abi.check(abi.lib.cgl_set_gemm_mode(mode))
    sims[mode] = MDStyleSim(algo, k, part_sizes=sizes)
    sims[mode].load(orc.net_g, orc.net_d)


def state(sim):
    out = {"D": sim.bank.rows().cpu().clone(), "Gt": sim.G.trunk.params[:, :sim.G.P_trunk].cpu().clone()}
    if sim.G.heads:
        out["Gh"] = sim.G.heads.params[:, :sim.G.P_head].cpu().clone()
    return out


def ref_state():
    out = {"D": torch.stack([torch.cat([p.detach().reshape(-1) for p in n.parameters()]) for n in orc.net_d]),
           "Gt": torch.stack([torch.cat([p.detach().reshape(-1) for p in g.model.parameters()]) for g in orc.net_g])}
    if hasattr(orc.net_g[0], "paths"):
        out["Gh"] = torch.stack([torch.cat([p.detach().reshape(-1) for p in path.parameters()])
                                 for g in orc.net_g for path in g.paths])
    return out


for r in range(rounds):
    real, n_real, z_d, z_g = _inputs(W, S, B, d, epoch, seed=50 + r)
    l_ref = orc.round(real, n_real, z_d, z_g)
    ls = {}
    for mode in (1, 0):
        abi.check(abi.lib.cgl_set_gemm_mode(mode))
        ls[mode] = sims[mode].round(real.cuda(), n_real.cuda(), z_d.cuda(), z_g.cuda()).cpu()
    ref = ref_state()
    print(f"round {r}: loss diff ffma {float((ls[1]-l_ref).abs().max()):.2e} auto {float((ls[0]-l_ref).abs().max()):.2e}")
    for mode, name in ((1, "ffma"), (0, "auto")):
        st = state(sims[mode])
        for key in st:
            a, b = st[key], ref[key]
            per_row = [(quantile_err(a[i], b[i], 0.9, 0.02), max_abs(a[i], b[i])) for i in range(a.shape[0])]
            print(f"   {name} {key}: " + " ".join(f"{q:.1e}/{m:.1e}" for q, m in per_row))
