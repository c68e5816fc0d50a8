"""torchrun entry of tests/test_gpu_nccl.py (WORLD_SIZE GPUs):
  * MD-GAN with its ONE server's clients dealt over the ranks (sim.MDSingleServerSim: replicated generator, all-gather of the
    losses, all-reduce of sum_i w_i dLoss_i/dXg) against the same rounds in one process (sim.MDStyleSim), incl. a
    discriminator swap and a group mean across the ranks;
  * FeGAN with the population dealt over the ranks (sim.FeGANSim: owners serve the group's members, all-reduce of the
    softmax(sk)-weighted partial sums) against the single-process FeGANSim;
  * FL-GAN rounds over a communicator (FLStyleSim + comm) against one process.
Prints `NCCL_ALGOS ok ...` on rank 0, or raises."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def gather_rows(t, world):
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t.contiguous())
    return torch.cat(out)


def check_mdgan(rank, local, world, dev, comm):
    from cgl_gan_b200 import models
    from cgl_gan_b200.sim import Knobs, MDSingleServerSim, MDStyleSim
    res = {}
    for shape, share in (((2,), "swap"), ((1, 28, 28), "group_mean")):
        d = 2 if shape == (2,) else 784
        W, B = 4 * world, 100
        sizes = [400 + 53 * i for i in range(W)]
        torch.manual_seed(21)
        k = Knobs(num_workers=W, num_servers=1, batch_size=B, img_shape=shape, E=2, d_share=share, num_communication=4)
        proto = MDStyleSim("mdgan", k, part_sizes=sizes, device=dev)
        g_mod = proto.G.make_module()
        d_mods = [models.Discriminator(shape) for _ in range(W)]
        sim = MDSingleServerSim("mdgan", k, part_sizes=sizes, device=dev, comm=comm, rank=rank, world=world)
        lo, hi = sim.lo, sim.hi
        sim.load([g_mod], d_mods[lo:hi])
        proto.load([g_mod], d_mods)
        gen = torch.Generator().manual_seed(5)
        for r in range(4):                       # shares at t = 4 and t = 2
            real = torch.tanh(torch.randn(W, B, d, generator=gen))
            z_d, z_g = torch.randn(1, B, 100, generator=gen), torch.randn(1, B, 100, generator=gen)
            l = sim.round(real[lo:hi].to(dev), None, z_d.to(dev), z_g.to(dev))
            l_ref = proto.round(real.to(dev), None, z_d.to(dev), z_g.to(dev))
            assert (l - l_ref).abs().max().item() < 1e-5, ("mdgan losses", shape, r)
        torch.cuda.synchronize()
        rows = gather_rows(sim.bank.rows(), world)
        e_d = rel(rows, proto.bank.rows())
        e_g = rel(sim.G.trunk.params, proto.G.trunk.params)
        # the all-reduce adds the ranks' partial sums of w_i dLoss_i/dXg in another order than the one-process client loop
        assert e_d < 5e-5 and e_g < 5e-5, (shape, e_d, e_g)
        # replicas stay bit-identical
        g_all = gather_rows(sim.G.trunk.params, world).view(world, *sim.G.trunk.params.shape)
        assert all(torch.equal(g_all[0], g_all[i]) for i in range(world)), "generator replicas diverged"
        res[share] = (e_d, e_g)
    return res


def check_fegan(rank, local, world, dev, comm):
    from cgl_gan_b200 import models
    from cgl_gan_b200.sim import FeGANSim, Knobs
    shape, d, B = (2,), 2, 100
    C = 3 * world
    torch.manual_seed(33)
    sk = [0.1 * (i % 7) + 0.05 for i in range(C)]
    groups = [[0, C - 1, 2], [1, 3, C - 2, 0], [2, C - 1]]
    k = Knobs(num_workers=C, num_servers=1, batch_size=B, img_shape=shape)
    one = FeGANSim(k, sk, groups, device=dev)
    g_mods = [one.G.make_module() for _ in range(C)]
    d_mods = [models.Discriminator(shape) for _ in range(C)]
    srv_g, srv_d = one.G.make_module(), models.Discriminator(shape)
    one.G.load_modules(g_mods)
    one.bank.load_modules(d_mods)
    one.load_global(srv_g, srv_d)
    sh = FeGANSim(k, sk, groups, device=dev, comm=comm, rank=rank, world=world)
    sh.G.load_modules(g_mods[sh.lo:sh.hi])
    sh.bank.load_modules(d_mods[sh.lo:sh.hi])
    sh.load_global(srv_g, srv_d)
    gen = torch.Generator().manual_seed(8)
    for r in range(len(groups)):
        group = groups[r]
        n = len(group)
        real = torch.tanh(torch.randn(n, B, d, generator=gen))
        z_d, z_g = torch.randn(n, B, 100, generator=gen), torch.randn(n, B, 100, generator=gen)
        g1, ids1 = one.begin_round()
        one.local_minibatch(real.to(dev), None, z_d.to(dev), z_g.to(dev), client_ids=ids1)
        one.end_round(g1, ids1)
        mine, ids = sh.begin_round()
        sel = [group.index(c) for c in mine]
        if mine:
            sh.local_minibatch(real[sel].to(dev), None, z_d[sel].to(dev), z_g[sel].to(dev), client_ids=ids)
        sh.end_round(mine, ids)
    torch.cuda.synchronize()
    e_d, e_g = rel(sh.p_d, one.p_d), rel(sh.p_g, one.p_g)
    assert e_d < 2e-6 and e_g < 2e-6, (e_d, e_g)
    rows = gather_rows(sh.bank.rows(), world)
    assert rel(rows, one.bank.rows()) < 2e-6
    return e_d, e_g


def check_flgan(rank, local, world, dev, comm):
    from cgl_gan_b200 import models
    from cgl_gan_b200.sim import FLStyleSim, Knobs
    shape, d, B = (2,), 2, 100
    Cl = 2
    C = Cl * world
    torch.manual_seed(44)
    one = FLStyleSim(Knobs(num_workers=C, num_servers=1, batch_size=B, img_shape=shape), device=dev)
    g_mod, d_mod = one.G.make_module(), models.Discriminator(shape)
    one.load_global(g_mod, d_mod)
    sh = FLStyleSim(Knobs(num_workers=Cl, num_servers=1, batch_size=B, img_shape=shape), device=dev, comm=comm)
    sh.load_global(g_mod, d_mod)
    gen = torch.Generator().manual_seed(9)
    lo = rank * Cl
    for r in range(2):
        real = torch.tanh(torch.randn(C, B, d, generator=gen))
        z_d, z_g = torch.randn(C, B, 100, generator=gen), torch.randn(C, B, 100, generator=gen)
        one.local_minibatch(real.to(dev), None, z_d.to(dev), z_g.to(dev))
        one.aggregate()
        sh.local_minibatch(real[lo:lo + Cl].to(dev), None, z_d[lo:lo + Cl].to(dev), z_g[lo:lo + Cl].to(dev))
        sh.aggregate()
    torch.cuda.synchronize()
    e_d = rel(sh.bank.rows()[0], one.bank.rows()[0])
    e_g = rel(sh.G.trunk.params[0], one.G.trunk.params[0])
    assert e_d < 2e-6 and e_g < 2e-6, (e_d, e_g)
    return e_d, e_g


def check_graph_rounds(rank, local, world, dev, comm):
    """MDStyleSim.round_graph over a communicator (the cloud all-reduce is captured with the round) == eager rounds."""
    from cgl_gan_b200 import models
    from cgl_gan_b200.sim import Knobs, MDStyleSim
    shape, d, B = (2,), 2, 100
    W, S = 4, 2                                  # per rank
    torch.manual_seed(70 + rank)
    k = Knobs(num_workers=W, num_servers=S, batch_size=B, segema=0.25, iid=1, img_shape=shape)
    sizes = [300 + 17 * i for i in range(W)]
    kw = dict(part_sizes=sizes, device=dev, comm=comm, server_offset=rank * S, total_data_len=sum(sizes) * world)
    a, b = MDStyleSim("cglgan", k, **kw), MDStyleSim("cglgan", k, **kw)
    g_mods = [a.G.make_module() for _ in range(S)]
    d_mods = [models.Discriminator(shape) for _ in range(W)]
    a.load(g_mods, d_mods)
    b.load(g_mods, d_mods)
    gen = torch.Generator().manual_seed(90 + rank)
    for r in range(5):
        real = torch.tanh(torch.randn(W, B, d, generator=gen)).to(dev)
        z_d, z_g = torch.randn(S, B, 100, generator=gen).to(dev), torch.randn(S, B, 100, generator=gen).to(dev)
        la = a.round(real, None, z_d, z_g)
        lb = b.round_graph(real, None, z_d, z_g)
        assert torch.equal(la, lb), ("graph round differs", r)
    torch.cuda.synchronize()
    assert torch.equal(a.bank.params, b.bank.params) and torch.equal(a.G.trunk.params, b.G.trunk.params)
    return True


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cgl_gan_b200.dist import ShardComm
    dev = f"cuda:{local}"
    comm = ShardComm()
    md = check_mdgan(rank, local, world, dev, comm)
    fe = check_fegan(rank, local, world, dev, comm)
    fl = check_flgan(rank, local, world, dev, comm)
    gr = check_graph_rounds(rank, local, world, dev, comm)
    if rank == 0:
        print(f"NCCL_ALGOS ok world={world} mdgan={md} fegan={fe} flgan={fl} graph={gr}", flush=True)
    comm.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
