"""torchrun entry of tests/test_gpu_nccl.py: a CGLGAN round sharded over WORLD_SIZE GPUs (servers dealt in
contiguous blocks, the cloud FedAvg through cgl_mix_allreduce over NCCL) against the same round on rank 0 alone.
Prints one line `NCCL_CHECK ok ...` per rank, or raises."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from cgl_gan_b200 import models
    from cgl_gan_b200.dist import ShardComm, shard_plan
    from cgl_gan_b200.sim import Knobs, MDStyleSim
    shape, B, d = (2,), 100, 2
    W, S = 8 * world, 4 * world          # 2 clients per server
    N = W // S
    sizes = [500 + 61 * i for i in range(W)]
    plan = shard_plan(W, S, world)
    slo, shi, clo, chi = plan[rank]
    torch.manual_seed(11)                 # every rank draws the same modules and inputs, keeps its shard
    kall = Knobs(num_workers=W, num_servers=S, batch_size=B, epoch=1, cloud_epoch=1, segema=0.25, iid=1, img_shape=shape)
    proto = MDStyleSim("cglgan", Knobs(num_workers=N, num_servers=1, batch_size=B, iid=1, img_shape=shape),
                       device=f"cuda:{local}")
    g_mods = [proto.G.make_module() for _ in range(S)]
    d_mods = [models.Discriminator(shape) for _ in range(W)]
    ref = MDStyleSim("cglgan", kall, part_sizes=sizes, device=f"cuda:{local}") if rank == 0 else None
    gen = torch.Generator().manual_seed(3)
    rounds = 3
    real = torch.tanh(torch.randn(rounds, W, B, d, generator=gen))
    z_d, z_g = torch.randn(rounds, S, B, 100, generator=gen), torch.randn(rounds, S, B, 100, generator=gen)

    comm = ShardComm()
    kloc = Knobs(num_workers=chi - clo, num_servers=shi - slo, batch_size=B, epoch=1, cloud_epoch=1, segema=0.25, iid=1,
                 img_shape=shape)
    sim = MDStyleSim("cglgan", kloc, part_sizes=sizes[clo:chi], device=f"cuda:{local}", comm=comm, server_offset=slo,
                     total_data_len=sum(sizes))
    sim.load(g_mods[slo:shi], d_mods[clo:chi])
    if rank == 0:
        ref.load(g_mods, d_mods)
    for r in range(rounds):
        l = sim.round(real[r, clo:chi].cuda(), None, z_d[r, slo:shi].cuda(), z_g[r, slo:shi].cuda())
        if rank == 0:
            l_ref = ref.round(real[r].cuda(), None, z_d[r].cuda(), z_g[r].cuda())
            assert (l - l_ref[slo:shi]).abs().max().item() < 1e-5, "losses of the sharded round differ"
    torch.cuda.synchronize()
    # gather every rank's trunk + discriminator rows on rank 0 and compare with the unsharded simulation
    trunk = sim.G.trunk.params.detach().contiguous()
    rows = sim.bank.rows().contiguous()
    tl = [torch.empty_like(trunk) for _ in range(world)]
    rl = [torch.empty_like(rows) for _ in range(world)]
    dist.all_gather(tl, trunk)            # equal shards here (W, S multiples of world)
    dist.all_gather(rl, rows)
    if rank == 0:
        t_all, r_all = torch.cat(tl), torch.cat(rl)
        t_ref, r_ref = ref.G.trunk.params.detach(), ref.bank.rows()
        et = ((t_all - t_ref).abs().max() / t_ref.abs().max()).item()
        er = ((r_all - r_ref).abs().max() / r_ref.abs().max()).item()
        # the all-reduce adds the ranks' partial sums in another order than the single-process row loop: ~1e-7
        assert et < 2e-6 and er < 2e-6, (et, er)
        print(f"NCCL_CHECK ok world={world} trunk_err={et:.2e} d_err={er:.2e}", flush=True)
    comm.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
