#!/usr/bin/env python
"""bench.py -- client GAN steps/sec of the CGLGAN MNIST round on N B200s (one process per GPU).

A step = one communication round of the CGLGAN MNIST simulation (BASELINE.json configs[1], scaled to
1024 clients per GPU as SURVEY.md 8d(5) prescribes): 256 edge servers x 4 clients per GPU, batch 100,
epoch 1, cloud_epoch 1, multi-head BN-MLP generator (4 heads / server), D = 784-512-256-1, BCE, Adam.
Every round: both generator passes, every client's D step (fwd real|fake, BCE, bwd, Adam), every
client's G-loss + dLoss/dXg, the server weighting + generator backward + Adam, and the cloud FedAvg of
the trunks (an NCCL all-reduce when N > 1).  value = clients * rounds / seconds, whole job.

  python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
  python bench.py --impl reference ...                     the reference's own CPU path (oracle restatement
                                                           of its PyTorch step, all host threads), bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "client_gan_steps_per_sec"
UNIT = "client-steps/s"
# algorithmic work of one MD-style MNIST client step
BYTES_PER_CLIENT_STEP = 24 * 533505 + 4 * 100 * 784 + 4 * 100 * 784 + 8 * 100 * 784  # 13.68 MB (own fake chunk)
FLOPS_PER_CLIENT_STEP = 2 * 100 * (8 * 532736 - 2 * 401408)                            # 0.692 GFLOP


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clients", type=int, default=1024, help="clients per GPU")
    ap.add_argument("--clients-per-server", type=int, default=4)
    ap.add_argument("--dataset", default="mnist", choices=["mnist", "2dmg"])
    ap.add_argument("--algo", default="cglgan", help="cglgan | capgan | mixed | mdgan | acgan (MD-style round) | flgan (FL-style step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-clients", type=int, default=16)
    return ap.parse_args()


def workload_config(args, world):
    shape = (1, 28, 28) if args.dataset == "mnist" else (2,)
    if args.algo == "flgan":
        return {
            "workload": f"FLGAN {args.dataset.upper()} step: {args.clients} clients/GPU, one local minibatch (D step + G step, "
                        f"batch 100) on every client, then the FedAvg of all G and D (FLGAN/MNIST/flgan.py:143-163,245-270)",
            "algo": "flgan", "dataset": f"synthetic {args.dataset}-shaped tanh(N(0,1))", "img_shape": list(shape),
            "num_workers": args.clients * world, "num_servers": 1, "batch_size": 100,
            "parallelism": f"clients sharded x{world}",
            "l2_policy": "inputs larger than L2 (>= 6 GB of per-client state streamed per step), no explicit flush",
        }
    return {
        "workload": f"{args.algo.upper()} {args.dataset.upper()} round: {args.clients} clients/GPU, "
                    f"{args.clients // args.clients_per_server} servers/GPU, batch 100, epoch 1, cloud_epoch 1, iid 1",
        "algo": args.algo, "dataset": f"synthetic {args.dataset}-shaped tanh(N(0,1))", "img_shape": list(shape),
        "num_workers": args.clients * world, "num_servers": args.clients // args.clients_per_server * world,
        "batch_size": 100, "epoch": 1, "cloud_epoch": 1, "iid": 1, "parallelism": f"clients sharded x{world}",
        "l2_policy": "inputs larger than L2 (>= 6 GB of per-client state streamed per round), no explicit flush",
    }


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference round, timed on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_round_rate(args, sample_clients, rounds, warmup):
    import torch
    from oracle.rounds import OracleMD
    torch.set_num_threads(os.cpu_count())
    shape = (1, 28, 28) if args.dataset == "mnist" else (2,)
    d = 784 if args.dataset == "mnist" else 2
    per = args.clients_per_server
    W, S, B = sample_clients, sample_clients // per, 100
    torch.manual_seed(20211212)
    if args.algo == "flgan":
        from oracle.rounds import OracleFL
        orc = OracleFL(W, B, shape)
        orc.load_global()
        g = torch.Generator().manual_seed(1)
        real = torch.tanh(torch.randn(W, B, d, generator=g))
        n_real = torch.full((W,), B, dtype=torch.int32)
        times = []
        for r in range(warmup + rounds):
            z_d, z_g = torch.randn(W, B, 100, generator=g), torch.randn(W, B, 100, generator=g)
            t0 = time.perf_counter()
            orc.local_minibatch(real, n_real, z_d, z_g)
            orc.aggregate()
            if r >= warmup:
                times.append(time.perf_counter() - t0)
        total = sum(times)
        return W * len(times) / total, total / len(times)
    orc = OracleMD(args.algo, W, S, B, shape, iid=1)
    g = torch.Generator().manual_seed(1)
    real = torch.tanh(torch.randn(1, W, B, d, generator=g))
    n_real = torch.full((1, W), B, dtype=torch.int32)
    times = []
    for r in range(warmup + rounds):
        z_d, z_g = torch.randn(S, B, 100, generator=g), torch.randn(S, B, 100, generator=g)
        t0 = time.perf_counter()
        orc.round(real, n_real, z_d, z_g)
        if r >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return W * len(times) / total, total / len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", args.gpus))
    sample = args.cpu_sample_clients
    rate, sec = cpu_round_rate(args, sample, args.steps, args.warmup)
    cfg = workload_config(args, max(world, 1))
    line = {
        "metric": METRIC, "value": rate, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "sample": f"{sample} clients / {sample // args.clients_per_server} servers of the same round, "
                                   f"{args.steps} rounds after {args.warmup} warm-up; torch CPU fp32, all host threads"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks sampling
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if rank == 0:
        ge.build()
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist.barrier()
    if rank != 0:
        ge.build()
    from cgl_gan_b200 import abi, models
    from cgl_gan_b200.dist import ShardComm
    from cgl_gan_b200.sim import Knobs, MDStyleSim
    abi.require_device()
    dev = torch.device("cuda", local_rank)
    comm = ShardComm() if world > 1 else None

    shape = (1, 28, 28) if args.dataset == "mnist" else (2,)
    d = 784 if args.dataset == "mnist" else 2
    C, per, B = args.clients, args.clients_per_server, 100
    S = C // per
    torch.manual_seed(20211212 + rank)
    fl = args.algo == "flgan"
    if fl:
        from cgl_gan_b200.sim import FLStyleSim
        k = Knobs(num_workers=C, num_servers=1, batch_size=B, img_shape=shape)
        sim = FLStyleSim(k, device=dev, comm=comm)
        sim.profile = False
        g_proto = [sim.G.make_module() for _ in range(4)]
        d_proto = [models.Discriminator(shape) for _ in range(8)]
        sim.G.load_modules([g_proto[c % 4] for c in range(C)])
        sim.bank.load_modules([d_proto[c % 8] for c in range(C)])

        def fl_step(real, n_real):
            """one local minibatch on every client, then Server.run's average (flgan.py:143-163)"""
            d_loss, g_loss = sim.local_minibatch(real, n_real)
            sim.aggregate()
            return g_loss
        sim.round = fl_step
        sim.client_step_ms = lambda: None
    else:
        k = Knobs(num_workers=C, num_servers=S, batch_size=B, epoch=1, cloud_epoch=1, segema=0.0, iid=1, img_shape=shape)
        sizes = [3000] * C
        sim = MDStyleSim(args.algo, k, part_sizes=sizes, device=dev, comm=comm, server_offset=rank * S,
                         total_data_len=3000 * C * world)
        # random-init weights of the reference architectures (torch default init), a few distinct modules tiled
        g_proto = [sim.G.make_module() for _ in range(4)]
        d_arch = abi.ARCH_D_2D if d == 2 else (abi.ARCH_D_MNIST2 if sim.loss_kind == abi.LOSS_CE else abi.ARCH_D_MNIST1)
        d_proto = [models.Discriminator(shape, arch=d_arch) for _ in range(8)]
        sim.G.load_modules([g_proto[s % 4] for s in range(S)])
        sim.bank.load_modules([d_proto[c % 8] for c in range(C)])

    # synthetic MNIST-shaped batches: a ring of pinned host buffers (e2e) and device-resident copies (value)
    ring = 2
    gen = torch.Generator().manual_seed(1234 + rank)
    host = [torch.tanh(torch.randn(C, B, d, generator=gen)).pin_memory() for _ in range(ring)]
    dev_real = [h.to(dev) for h in host]
    n_real_host = torch.full((C,), B, dtype=torch.int32).pin_memory()
    n_real_dev = n_real_host.to(dev)
    stage = [torch.empty(C, B, d, device=dev) for _ in range(2)]
    loss_host = (torch.empty(C, dtype=torch.float32) if fl else torch.empty(S, per, dtype=torch.float32)).pin_memory()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def resident_round(i):
        return sim.round(dev_real[i % ring], n_real_dev)

    copy_stream = torch.cuda.Stream()

    def e2e_round(i, pending):
        """H2D of this round's real batches from pinned memory (prefetched one round ahead on a copy stream),
        the round, D2H of the clients' G losses."""
        cur = torch.cuda.current_stream()
        buf, ev = pending
        cur.wait_event(ev)
        nxt = stage[(i + 1) % 2]
        with torch.cuda.stream(copy_stream):   # stage[(i+1)%2] is free: round i-1 was synchronised
            nxt.copy_(host[(i + 1) % ring], non_blocking=True)
            ev2 = torch.cuda.Event()
            ev2.record(copy_stream)
        loss = sim.round(buf, n_real_dev)
        loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()     # the user reads the round's losses
        return (nxt, ev2)

    launch0 = [0]

    def timed(fn, steps, warmup, e2e=False, profile=False):
        pending = None
        if e2e:
            with torch.cuda.stream(copy_stream):
                stage[0].copy_(host[0], non_blocking=True)
                ev = torch.cuda.Event(); ev.record(copy_stream)
            pending = (stage[0], ev)
        for i in range(warmup):
            pending = fn(i, pending) if e2e else fn(i)
        sync_all()
        if profile:                        # (re)start the per-kernel event pairs: the timed rounds only
            sim.profile = True
            abi.profile_enable(True)
        launch0[0] = abi.launch_count()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for i in range(steps):
            pending = fn(warmup + i, pending) if e2e else fn(warmup + i)
        stop.record()
        sync_all()
        ms = torch.tensor([start.elapsed_time(stop)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # CUDA-event pairs around every engine kernel, on the launching stream, over the timed rounds
    ms = timed(resident_round, args.steps, args.warmup, profile=True)
    launches = abi.launch_count() - launch0[0]
    client_ms = sim.client_step_ms()
    sim.profile = False
    kernels = abi.profile_summary()
    abi.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = timed(e2e_round, args.steps, args.warmup, e2e=True)

    total_clients = C * world
    value = total_clients * args.steps / (ms / 1e3)
    e2e_value = total_clients * args.steps / (ms_e2e / 1e3)

    peaks = {}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peaks = json.load(open(p))
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md: 6650 GB/s, 1.4 PF sustained)"
    rounds_profiled = args.steps
    table = {}
    for name, k in kernels.items():
        sec = k["ms"] / 1e3
        table[name] = {"ms_per_round": k["ms"] / rounds_profiled, "launches_per_round": k["launches"] / rounds_profiled,
                       "ms_per_launch": k["ms"] / k["launches"],
                       "algorithmic_GBps": (k["bytes"] / sec / 1e9) if sec > 0 else None,
                       "algorithmic_TFLOPps": (k["flops"] / sec / 1e12) if sec > 0 else None,
                       "bytes_per_launch": k["bytes"] / k["launches"], "flops_per_launch": k["flops"] / k["launches"]}
    traffic_tab = {}
    tpath = os.path.join(ROOT, "profiles", "traffic_r1.json")   # dram bytes per launch, from the ncu capture of a round
    if os.path.exists(tpath):
        traffic_tab = json.load(open(tpath)).get(args.dataset, {})

    def roofline_of(name):
        dk = table[name]
        hbm_bound = "[tcgen05]" not in name and "[ffma]" not in name or "wgrad+adam" in name
        if hbm_bound:
            r = {"kernel": name, "bound": "hbm", "achieved": dk["algorithmic_GBps"], "peak": hbm_peak, "unit": "GB/s",
                 "frac": dk["algorithmic_GBps"] / hbm_peak}
        elif "[tcgen05]" in name:
            # 3xTF32: every fp32 multiply-add is three tf32 MMAs, and tf32 runs at half the bf16 rate, so the tensor
            # pipe does 6 bf16-rate units of work per algorithmic fp32 FLOP; the peak is the measured dense bf16 rate
            issued = 6.0 * dk["algorithmic_TFLOPps"]
            r = {"kernel": name, "bound": "tensor", "achieved": issued, "peak": tf_peak, "unit": "TFLOP/s",
                 "frac": issued / tf_peak, "fp32_equivalent_TFLOPps": dk["algorithmic_TFLOPps"],
                 "note": "achieved = 6 x fp32-equivalent FLOP/s: 3 tf32 MMAs per product at half the bf16 rate (the work the "
                         "tensor pipe executes for strict-fp32 results), against the measured dense bf16 peak"}
        else:
            r = {"kernel": name, "bound": "fp32 FFMA", "achieved": dk["algorithmic_TFLOPps"], "peak": 75.0, "unit": "TFLOP/s",
                 "frac": dk["algorithmic_TFLOPps"] / 75.0, "note": "exact-fp32 FFMA kernel; peak = 148 SMs x 128 FMA/clk x 1.965 GHz"}
        r.update({"traffic": traffic_tab.get(name, {}).get("dram_bytes_per_launch"),
                  "traffic_source": "profiles/traffic_r1.json (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch of this class)"
                                    if name in traffic_tab else None,
                  "peak_source": peak_src, "ms_per_launch": dk["ms_per_launch"],
                  "share_of_round": dk["ms_per_round"] / (ms / args.steps),
                  "bytes_per_launch": dk["bytes_per_launch"], "flops_per_launch": dk["flops_per_launch"],
                  "timing": "cudaEvent pairs on the launching stream around every launch of this kernel class, "
                            "over the timed rounds of this run"})
        return r

    order = sorted(table, key=lambda n: -table[n]["ms_per_round"])
    # the two GEMM classes at the top are within a per-cent of each other and swap places from run to run: within 3 %
    # the HBM-bound weight-gradient + Adam kernel (the bound SURVEY.md 8d names for the client step) is reported first
    if len(order) > 1 and "wgrad+adam" in order[1] and \
            table[order[1]]["ms_per_round"] > 0.97 * table[order[0]]["ms_per_round"]:
        order[0], order[1] = order[1], order[0]
    roofline = roofline_of(order[0]) if order else None
    roofline_2 = roofline_of(order[1]) if len(order) > 1 else None
    client = None
    if client_ms and d == 784:
        client = {"client_step_ms_per_round": client_ms, "algorithmic_bytes_per_client_step": BYTES_PER_CLIENT_STEP,
                  "hbm_GBps": BYTES_PER_CLIENT_STEP * C / (client_ms / 1e3) / 1e9,
                  "hbm_frac": BYTES_PER_CLIENT_STEP * C / (client_ms / 1e3) / 1e9 / hbm_peak,
                  "fp32_TFLOPps": FLOPS_PER_CLIENT_STEP * C / (client_ms / 1e3) / 1e12}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "round_time_s": ms / args.steps / 1e3,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(C * B * d * 4 * world), "d2h_bytes_per_step": int(S * per * 4 * world)},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_second_kernel": roofline_2,
            "client_step": client,
            "kernels": table,
        }
        if not args.no_cpu_baseline and world == 1:
            sample = args.cpu_sample_clients
            rate, sec = cpu_round_rate(args, sample, 20, 3)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{sample} clients / {sample // per} servers of the same round, 20 rounds "
                                              f"after 3 warm-up ({sec * 1e3:.1f} ms/round); torch CPU fp32, all host threads"}
        print(json.dumps(line))
    if comm is not None:
        comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
