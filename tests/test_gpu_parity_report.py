"""Measured parity, written down: for every algorithm x dataset x GEMM mode, after 1 and after 10 rounds, the distance
between the engine and the oracle for the discriminators, the generators, dLoss/dXg on a probe batch and the losses --
max-norm, q90, q99, fraction of elements outside 1e-5, relative L2 -- next to the oracle's OWN distance between a run with
all host threads and a run with one thread on inputs moved by ONE ULP (another fp32 summation order of the same reference
code, and the sensitivity of the dynamics to a perturbation of the size of a single rounding: some shapes run bit-identically
on 1 and N threads, and their self-distance would otherwise read zero).

The numbers go to profiles/parity_r2.json (CGL_PARITY_OUT overrides the path; the GPU box writes gpurun_out/parity_r2.json,
which is copied into profiles/). The assertions are stated against K x that self-noise, not against a fixed lenient
constant: an Adam step is ill-conditioned where a gradient cancels to below eps = 1e-8 or a pre-activation sits on the
LeakyReLU kink, and how often that happens in a given run is exactly what the reference-vs-itself distance measures."""
import copy
import json
import os

import pytest
import torch

from helpers import bn_fed_biases
from oracle import steps as osteps
from oracle.rounds import OracleFL, OracleMD

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.environ.get("CGL_PARITY_OUT") or os.path.join(ROOT, "profiles", "parity_r2.json")
LR = 2e-4
FLOOR = 100 * LR      # tensors that start at zero are measured against the scale of a weight tensor (helpers.py)
K_NOISE = 5.0         # engine-vs-oracle may be this many times the oracle's own distance from itself

CASES = [
    # id, algo, img_shape, workers, servers, iid, segema, knobs
    ("cglgan_2dmg", "cglgan", (2,), 10, 5, 1, 0.0, {}),
    ("cglgan_2dmg_iid0", "cglgan", (2,), 4, 2, 0, 0.0, {}),
    ("cglgan_mnist", "cglgan", (1, 28, 28), 8, 2, 1, 0.3, {}),
    ("capgan_mnist", "capgan", (1, 28, 28), 4, 2, 1, 0.0, {}),
    ("capgan_copy_mnist", "capgan_copy", (1, 28, 28), 4, 2, 1, 0.0, {}),
    ("mixed_mnist_E5", "mixed", (1, 28, 28), 4, 2, 1, 0.5, {"E": 5, "d_share": "group_mean"}),
    ("mdgan_mnist", "mdgan", (1, 28, 28), 3, 1, 1, 0.0, {}),
    ("mdgan_2dmg_swapE2", "mdgan", (2,), 4, 1, 1, 0.0, {"E": 2, "d_share": "swap"}),
    ("acgan_mnist_E5", "acgan", (1, 28, 28), 4, 2, 1, 0.0, {"E": 5, "d_share": "group_mean"}),
]
ROUNDS = (1, 10)


def metrics(a, b, floor=FLOOR):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    scale = max(b.abs().max().item(), floor)
    e = (a - b).abs() / scale
    n = e.numel()
    srt = e.sort().values
    return {"max": e.max().item(), "q90": srt[min(n - 1, int(0.9 * n))].item(), "q99": srt[min(n - 1, int(0.99 * n))].item(),
            "frac_gt_1e-5": (e > 1e-5).double().mean().item(),
            "rel_l2": ((a - b).norm() / max(b.norm().item(), floor * n ** 0.5)).item(), "numel": n, "scale": scale}


def worst(ms):
    out = {k: max(m[k] for m in ms) for k in ("max", "q90", "q99", "frac_gt_1e-5", "rel_l2")}
    out["tensors"] = len(ms)
    return out


def g_tensors(net, skip_noise_fed=True):
    """Generator parameters by name; Linear biases that feed a BatchNorm have an identically-zero true gradient (Adam
    amplifies rounding noise there) and are reported separately."""
    skip = bn_fed_biases(net) if skip_noise_fed else set()
    main, noise = {}, {}
    for k, v in net.state_dict().items():
        if not v.dim() or "running" in k:
            continue
        (noise if k in skip else main)[k] = v.detach().clone()
    return main, noise


def one_ulp(x, seed):
    """x moved by one unit in the last place (sign at random): the smallest perturbation fp32 can express."""
    g = torch.Generator().manual_seed(seed)
    sgn = torch.randint(0, 2, x.shape, generator=g).float() * 2 - 1
    return x + sgn * x.abs() * 2.0 ** -23


def _inputs(C, S, B, d, seed):
    g = torch.Generator().manual_seed(seed)
    real = torch.tanh(torch.randn(1, C, B, d, generator=g))
    n_real = torch.full((1, C), B, dtype=torch.int32)
    n_real[0, 0] = 41
    real[0, 0, 41:] = 0
    return real, n_real, torch.randn(S, B, 100, generator=g), torch.randn(S, B, 100, generator=g)


def _update_report(key, mode, entry):
    rep = {}
    if os.path.exists(OUT):
        try:
            rep = json.load(open(OUT))
        except Exception:
            rep = {}
    rep.setdefault("_about", "engine vs oracle (gpu) and oracle(all threads) vs oracle(1 thread) (self); errors are |a-b| / "
                             "max(max|ref|, 0.02) per tensor, worst tensor reported; tests/test_gpu_parity_report.py")
    rep.setdefault(key, {})[mode] = entry
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    json.dump(rep, open(OUT, "w"), indent=1, sort_keys=True)


@pytest.fixture(params=[1, 0], ids=["ffma", "auto"])
def gemm_mode(request, lib):
    lib.check(lib.lib.cgl_set_gemm_mode(request.param))
    yield "ffma" if request.param == 1 else "auto"
    lib.check(lib.lib.cgl_set_gemm_mode(0))


# Bars, calibrated on profiles/parity_r2.json (B200, both GEMM modes, all cases above):
#   after ONE round the bulk of every tensor agrees to rounding: measured q90 <= 4.4e-7 (bar 1e-6, ten times below the
#   1e-5 of BASELINE.json), at most 1.6 % of the elements outside 1e-5 -- and that 1.6 % is the oracle's own
#   1-vs-N-thread figure for the same case;
#   ten FREE-RUNNING rounds are chaotic: one flipped ill-conditioned element moves everything downstream by ~lr per round.
#   The oracle against itself (one-ulp inputs) reaches q90 = 5.2e-4 there, the engine the same; whether a given run hits
#   such an event is luck, so the bar is the larger of K x this run's own noise and the worst noise of the table (1e-3).
#   The ten-round comparison that is NOT at the mercy of that luck is test_md_parity_teacher_forced below.
ONE_ROUND_Q90 = 1e-6
ONE_ROUND_FRAC = 2e-2
CHAOS_Q90 = 1e-3


def _check(tag, steps, gpu, noise, max_bound):
    """engine-vs-oracle against K x the oracle's own noise and the Adam bound (2.2 lr per step) on the max-norm."""
    if steps == 1:
        assert gpu["q90"] <= max(ONE_ROUND_Q90, K_NOISE * noise["q90"]), (tag, "q90", gpu, noise)
        assert gpu["frac_gt_1e-5"] <= max(ONE_ROUND_FRAC, 2 * noise["frac_gt_1e-5"]), (tag, "frac", gpu, noise)
        assert gpu["rel_l2"] <= max(1e-3, K_NOISE * noise["rel_l2"]), (tag, "rel_l2", gpu, noise)
    else:
        assert gpu["q90"] <= max(CHAOS_Q90, K_NOISE * noise["q90"]), (tag, "q90", gpu, noise)
        assert gpu["rel_l2"] <= max(2 * CHAOS_Q90, K_NOISE * noise["rel_l2"]), (tag, "rel_l2", gpu, noise)
    assert gpu["max"] <= max_bound, (tag, "max", gpu, max_bound)


@pytest.mark.parametrize("cid,algo,shape,W,S,iid,segema,extra", CASES, ids=[c[0] for c in CASES])
def test_md_parity_measured(lib, gemm_mode, cid, algo, shape, W, S, iid, segema, extra):
    from cgl_gan_b200.sim import Knobs, MDStyleSim
    torch.manual_seed(20211212)
    B = 100
    d = 1
    for s in shape:
        d *= s
    sizes = [1000 + 137 * i for i in range(W)]
    nc = 12      # the down-counter of the reference: shares / cloud rounds fall inside the 10 rounds
    orc = OracleMD(algo, W, S, B, shape, iid=iid, part_sizes=sizes, segema=segema, weights_init=(algo == "mixed"),
                   num_communication=nc, **extra)
    k = Knobs(num_workers=W, num_servers=S, batch_size=B, epoch=1, segema=segema, iid=iid, img_shape=shape,
              num_communication=nc, **extra)
    sim = MDStyleSim(algo, k, part_sizes=sizes)
    sim.load(orc.net_g, orc.net_d)
    orc1 = copy.deepcopy(orc)
    threads = torch.get_num_threads()
    entry = {}
    loss_err, loss_self = 0.0, 0.0
    gp = torch.Generator().manual_seed(999)
    probe = torch.tanh(torch.randn(W, B, d, generator=gp) * 0.5)
    for r in range(max(ROUNDS)):
        real, n_real, z_d, z_g = _inputs(W, S, B, d, seed=50 + r)
        l_ref = orc.round(real, n_real, z_d, z_g)
        torch.set_num_threads(1)
        try:
            l_ref1 = orc1.round(one_ulp(real, 7000 + r), n_real, one_ulp(z_d, 8000 + r), one_ulp(z_g, 9000 + r))
        finally:
            torch.set_num_threads(threads)
        l_gpu = sim.round(real.cuda(), n_real.cuda(), z_d.cuda(), z_g.cuda()).cpu()
        loss_err = max(loss_err, (l_gpu - l_ref).abs().max().item())
        loss_self = max(loss_self, (l_ref1 - l_ref).abs().max().item())
        if r + 1 not in ROUNDS:
            continue
        # discriminators
        dm, dn = [], []
        for c in range(W):
            ref = torch.cat([p.detach().reshape(-1) for p in orc.net_d[c].parameters()])
            ref1 = torch.cat([p.detach().reshape(-1) for p in orc1.net_d[c].parameters()])
            dm.append(metrics(sim.bank.rows()[c], ref))
            dn.append(metrics(ref1, ref))
        # generators
        gm, gn, bm, bnn = [], [], [], []
        for s in range(S):
            m = sim.G.make_module()
            sim.G.store_module(s, m)
            got, got_b = g_tensors(m)
            ref, ref_b = g_tensors(orc.net_g[s])
            ref1, ref1_b = g_tensors(orc1.net_g[s])
            for key in ref:
                gm.append(metrics(got[key], ref[key]))
                gn.append(metrics(ref1[key], ref[key]))
            for key in ref_b:
                bm.append(metrics(got_b[key], ref_b[key]))
                bnn.append(metrics(ref1_b[key], ref_b[key]))
        # dLoss/dXg of a probe batch through the current discriminators
        _, dx = sim.bank.g_loss_raw(probe.cuda())
        crit = osteps.make_loss(orc.kind)
        xm, xn = [], []
        for c in range(W):
            outs = []
            for net in (orc.net_d[c], orc1.net_d[c]):
                x = probe[c].clone().requires_grad_(True)
                osteps.worker_g_loss(net, crit, orc.kind, x, B).backward()
                outs.append(x.grad)
            xm.append(metrics(dx[c], outs[0], floor=0.0))
            xn.append(metrics(outs[1], outs[0], floor=0.0))
        rec = {"D": {"gpu": worst(dm), "self": worst(dn)}, "G": {"gpu": worst(gm), "self": worst(gn)},
               "dXg_probe": {"gpu": worst(xm), "self": worst(xn)},
               "loss_abs": {"gpu": loss_err, "self": loss_self}}
        if bm:
            rec["G_bias_before_batchnorm"] = {"gpu": worst(bm), "self": worst(bnn)}
        entry[f"rounds_{r + 1}"] = rec
    _update_report(cid, gemm_mode, entry)          # written before anything is asserted: a failing case is still on record
    for steps in ROUNDS:
        rec = entry[f"rounds_{steps}"]
        _check((cid, "D", steps), steps, rec["D"]["gpu"], rec["D"]["self"], 2.2 * LR * steps / FLOOR)
        _check((cid, "G", steps), steps, rec["G"]["gpu"], rec["G"]["self"], 2.2 * LR * steps / FLOOR)
        assert rec["loss_abs"]["gpu"] < 1e-4, (cid, steps, rec["loss_abs"])


@pytest.mark.parametrize("shape", [(2,), (1, 28, 28)], ids=["flgan_2dmg", "flgan_mnist"])
def test_fl_parity_measured(lib, gemm_mode, shape):
    from cgl_gan_b200.sim import FLStyleSim, Knobs
    torch.manual_seed(7)
    C, B = 3, 100
    d = 1
    for s in shape:
        d *= s
    orc = OracleFL(C, B, shape)
    orc.load_global()
    orc1 = copy.deepcopy(orc)
    sim = FLStyleSim(Knobs(num_workers=C, num_servers=1, batch_size=B, img_shape=shape))
    sim.load_global(orc.srv_g, orc.srv_d)
    g = torch.Generator().manual_seed(1)
    threads = torch.get_num_threads()
    entry = {}
    loss_err = 0.0
    for r in range(max(ROUNDS)):
        real = torch.tanh(torch.randn(C, B, d, generator=g))
        n_real = torch.tensor([B, 60, B], dtype=torch.int32)
        real[1, 60:] = 0
        z_d, z_g = torch.randn(C, B, 100, generator=g), torch.randn(C, B, 100, generator=g)
        dl_ref, gl_ref = orc.local_minibatch(real, n_real, z_d, z_g)
        torch.set_num_threads(1)
        try:
            orc1.local_minibatch(one_ulp(real, 7000 + r), n_real, one_ulp(z_d, 8000 + r), one_ulp(z_g, 9000 + r))
            orc1.aggregate()
        finally:
            torch.set_num_threads(threads)
        dl, gl = sim.local_minibatch(real.cuda(), n_real.cuda(), z_d.cuda(), z_g.cuda())
        loss_err = max(loss_err, (dl.cpu() - dl_ref).abs().max().item(), (gl.cpu() - gl_ref).abs().max().item())
        orc.aggregate()
        sim.aggregate()
        if r + 1 not in ROUNDS:
            continue
        ref = torch.cat([p.detach().reshape(-1) for p in orc.net_d[0].parameters()])
        ref1 = torch.cat([p.detach().reshape(-1) for p in orc1.net_d[0].parameters()])
        m = sim.G.make_module()
        sim.G.store_module(0, m)
        got, _ = g_tensors(m)
        gref, _ = g_tensors(orc.net_g[0])
        gref1, _ = g_tensors(orc1.net_g[0])
        rec = {"D": {"gpu": worst([metrics(sim.bank.rows()[0], ref)]), "self": worst([metrics(ref1, ref)])},
               "G": {"gpu": worst([metrics(got[k], gref[k]) for k in gref]),
                     "self": worst([metrics(gref1[k], gref[k]) for k in gref])},
               "loss_abs": {"gpu": loss_err}}     # max over the rounds so far
        entry[f"rounds_{r + 1}"] = rec
    _update_report("flgan_" + ("2dmg" if d == 2 else "mnist"), gemm_mode, entry)
    for steps in ROUNDS:
        rec = entry[f"rounds_{steps}"]
        _check(("fl", "D", steps), steps, rec["D"]["gpu"], rec["D"]["self"], 2.2 * LR * steps / FLOOR)
        _check(("fl", "G", steps), steps, rec["G"]["gpu"], rec["G"]["self"], 2.2 * LR * steps / FLOOR)
        assert rec["loss_abs"]["gpu"] < 1e-4


# ---- ten rounds along the oracle's trajectory --------------------------------------------------------------------------
def _adam_flat(params, opt):
    ms, vs, step = [], [], 0
    for p in params:
        st = opt.state.get(p, {})
        if st:
            ms.append(st["exp_avg"].reshape(-1))
            vs.append(st["exp_avg_sq"].reshape(-1))
            step = int(st["step"])
        else:
            ms.append(torch.zeros(p.numel()))
            vs.append(torch.zeros(p.numel()))
    return torch.cat(ms), torch.cat(vs), step


def _load_bank(bank, nets_params, opts, stats=None):
    """rows of a packed bank <- (parameters, Adam moments, step[, BatchNorm running statistics]) of oracle modules"""
    R = len(nets_params)
    P = sum(p.numel() for p in nets_params[0])
    prm, mm, vv = torch.zeros(R, bank.params.shape[1]), torch.zeros(R, bank.params.shape[1]), torch.zeros(R, bank.params.shape[1])
    steps = torch.zeros(R, dtype=torch.int32)
    for r, (params, opt) in enumerate(zip(nets_params, opts)):
        prm[r, :P] = torch.cat([p.detach().reshape(-1) for p in params])
        m, v, t = _adam_flat(params, opt)
        mm[r, :P], vv[r, :P], steps[r] = m, v, t
    bank.params.copy_(prm)
    bank.adam_m.copy_(mm)
    bank.adam_v.copy_(vv)
    bank.step.copy_(steps)
    if stats is not None and bank.stats.shape[1] >= 1:
        st = torch.zeros(R, bank.stats.shape[1])
        for r, s_ in enumerate(stats):
            if s_.numel():
                st[r, :s_.numel()] = s_
        bank.stats.copy_(st)


def sync_engine_to_oracle(sim, orc):
    """Teacher forcing: the engine's whole state (every D and G row, Adam moments and step counters, BatchNorm running
    statistics, Lambda, the round counter) is overwritten with the oracle's."""
    from cgl_gan_b200.layout import flatten_bn_stats
    _load_bank(sim.bank, [list(n.parameters()) for n in orc.net_d], orc.opti_d)
    G = sim.G
    _load_bank(G.trunk, [list(n.model.parameters()) for n in orc.net_g], orc.opti_g,
               [flatten_bn_stats(n.model) for n in orc.net_g])
    if G.N:
        nets, opts, stats = [], [], []
        for s, n in enumerate(orc.net_g):
            for path in n.paths:
                nets.append(list(path.parameters()))
                opts.append(orc.opti_g[s])
                stats.append(flatten_bn_stats(path))
        _load_bank(G.heads, nets, opts, stats)
    sim.Lambda.copy_(torch.stack([L.detach().reshape(()) for L in orc.Lambda]))
    sim.t = orc.t


FORCED = [c for c in CASES if c[0] in ("cglgan_2dmg", "cglgan_mnist", "mixed_mnist_E5", "capgan_mnist", "mdgan_2dmg_swapE2")]


@pytest.mark.parametrize("cid,algo,shape,W,S,iid,segema,extra", FORCED, ids=[c[0] for c in FORCED])
def test_md_parity_teacher_forced(lib, gemm_mode, cid, algo, shape, W, S, iid, segema, extra):
    """Ten rounds along the ORACLE's trajectory: before every round the engine's state is overwritten with the oracle's,
    so each round is a one-round comparison from a realistic trained state (Adam moments with history, step counts 1..10 and
    their bias corrections, cloud mixes and discriminator shares on their rounds) and nothing is left to chaotic
    amplification. The one-round bar must hold at EVERY round, not only from the initial weights."""
    from cgl_gan_b200.sim import Knobs, MDStyleSim
    torch.manual_seed(20211212)
    B = 100
    d = 1
    for s in shape:
        d *= s
    sizes = [1000 + 137 * i for i in range(W)]
    nc = 12
    orc = OracleMD(algo, W, S, B, shape, iid=iid, part_sizes=sizes, segema=segema, weights_init=(algo == "mixed"),
                   num_communication=nc, **extra)
    k = Knobs(num_workers=W, num_servers=S, batch_size=B, epoch=1, segema=segema, iid=iid, img_shape=shape,
              num_communication=nc, **extra)
    sim = MDStyleSim(algo, k, part_sizes=sizes)
    sim.load(orc.net_g, orc.net_d)
    rounds = []
    threads = torch.get_num_threads()
    for r in range(10):
        sync_engine_to_oracle(sim, orc)
        twin = copy.deepcopy(orc)            # the oracle's one-ulp twin, from the SAME state: this round's own conditioning
        real, n_real, z_d, z_g = _inputs(W, S, B, d, seed=150 + r)
        l_ref = orc.round(real, n_real, z_d, z_g)
        torch.set_num_threads(1)
        try:
            twin.round(one_ulp(real, 7100 + r), n_real, one_ulp(z_d, 8100 + r), one_ulp(z_g, 9100 + r))
        finally:
            torch.set_num_threads(threads)
        l_gpu = sim.round(real.cuda(), n_real.cuda(), z_d.cuda(), z_g.cuda()).cpu()
        dm, dn = [], []
        for c in range(W):
            ref = torch.cat([p.detach().reshape(-1) for p in orc.net_d[c].parameters()])
            dm.append(metrics(sim.bank.rows()[c], ref))
            dn.append(metrics(torch.cat([p.detach().reshape(-1) for p in twin.net_d[c].parameters()]), ref))
        gm, gn = [], []
        for s in range(S):
            m = sim.G.make_module()
            sim.G.store_module(s, m)
            got, _ = g_tensors(m)
            ref, _ = g_tensors(orc.net_g[s])
            tw, _ = g_tensors(twin.net_g[s])
            gm += [metrics(got[key], ref[key]) for key in ref]
            gn += [metrics(tw[key], ref[key]) for key in ref]
        rounds.append({"D": worst(dm), "G": worst(gm), "D_self": worst(dn), "G_self": worst(gn),
                       "loss_abs": (l_gpu - l_ref).abs().max().item()})
    _update_report(cid, gemm_mode + "_teacher_forced", {"per_round": rounds})
    # Measured (profiles/parity_r2.json): in a quiet round engine and oracle agree to ONE ULP (max-norm 1.2e-7, q90 6e-8, no
    # element outside 1e-5). In other rounds a pre-activation of some sample sits within rounding of LeakyReLU's kink, the
    # two summation orders disagree on its side, and that sample's contribution moves every weight downstream by ~1e-3 of
    # its update: bursts of q90 4e-6 .. 2.4e-5. The oracle's one-ulp twin shows bursts of the same size (up to 2.0e-5) in
    # about a third of the rounds -- in OTHER rounds than the engine, so a round-by-round ratio is meaningless. Hence:
    #   every round : q90 <= 1e-4 (5 x the twin's largest burst) or K x the twin's q90 of that round, max-norm inside the
    #                 Adam bound, losses within 1e-5;
    #   typical round (median over the ten): the stated 1e-5;  and at least 4 of the 10 rounds at rounding level (1e-6).
    for what in ("D", "G"):
        q = sorted(rec[what]["q90"] for rec in rounds)
        assert 0.5 * (q[4] + q[5]) <= 1e-5, (cid, what, "median q90", q)
        assert sum(1 for v in q if v <= ONE_ROUND_Q90) >= 4, (cid, what, "rounds at rounding level", q)
    for r, rec in enumerate(rounds):
        for what in ("D", "G"):
            g, n = rec[what], rec[what + "_self"]
            assert g["q90"] <= max(1e-4, K_NOISE * n["q90"]), (cid, r, what, g, n)
            assert g["max"] <= 2.2 * LR / FLOOR, (cid, r, what, g)
        assert rec["loss_abs"] < 1e-5, (cid, r, rec["loss_abs"])
