#!/usr/bin/env python
"""bench.py -- client GAN steps/sec of the CGL-GAN simulated-client round on N B200s (one process per GPU).

HEADLINE LINE (BASELINE.json configs[1], scaled to 1024 clients per GPU as SURVEY.md 8d(5) prescribes): one communication
round of the CGLGAN MNIST simulation -- 256 edge servers x 4 clients per GPU, batch 100, epoch 1, cloud_epoch 1, multi-head
BN-MLP generator (4 heads / server), D = 784-512-256-1, BCE, Adam. Every round: both generator passes, every client's D step
(fwd real|fake, BCE, bwd, Adam), every client's G loss + dLoss/dXg, the server weighting + generator backward + Adam, and
the cloud FedAvg of the trunks (an NCCL all-reduce when N > 1).  value = clients * rounds / seconds, whole job, weak scaling.

The same JSON line carries, under "configs", the other workloads BASELINE.json names, each measured in the same run with its
own roofline and (N = 1) its own CPU baseline:
  strong_1024_mnist / strong_1024_2dmg : 1024 clients IN TOTAL dealt over the N GPUs (the north-star scaling sentence)
  2dmg_repo_default                    : CGLGAN 2DMG, 10 workers / 5 servers (configs[0]), CUDA-graph rounds
  capgan_mixg_E5                       : mixed-gan.py (CAPGAN + Mix-G), neighbour discriminator share every E = 5 rounds (configs[2])
  mdgan_single_server                  : MD-GAN, ONE server, its clients dealt over the GPUs (configs[3])
  flgan / fegan                        : FL-GAN local minibatch + FedAvg; FeGAN with frac_workers = 0.2 (configs[3])

  python bench.py --gpus N --steps K --warmup W            (torchrun for N > 1)
  python bench.py --impl reference ...                     the reference's own CPU path (oracle restatement of its PyTorch
                                                           step, all host threads), bounded sample
"""
import argparse
import gc
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
SUB_STEPS, SUB_WARMUP = 6, 3


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clients", type=int, default=1024, help="clients per GPU of the headline workload")
    ap.add_argument("--clients-per-server", type=int, default=4)
    ap.add_argument("--dataset", default="mnist", choices=["mnist", "2dmg"])
    ap.add_argument("--algo", default="cglgan", help="cglgan | capgan | mixed | mdgan | acgan (MD-style round) | flgan (FL-style step)")
    ap.add_argument("--configs", default="all", help="all | none | comma list of the sub-records to measure")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-clients", type=int, default=64)
    return ap.parse_args()


def shape_of(dataset):
    return ((1, 28, 28), 784) if dataset == "mnist" else ((2,), 2)


def workload_config(args, world):
    shape, _ = shape_of(args.dataset)
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
        "cloud_mode": "intended (the Cloud average of the trunks reaches the generators; the as-written scripts load nothing, "
                      "SURVEY 3.5.2 -- Knobs.cloud_mode)",
        "l2_policy": "inputs larger than L2 (>= 6 GB of per-client state streamed per round), no explicit flush",
    }


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference round, timed on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_round_rate(algo, dataset, per, sample_clients, rounds, warmup, knobs=None):
    """-> (client-steps/s, seconds per round) of the oracle on `sample_clients` clients with all host threads."""
    import torch
    torch.set_num_threads(os.cpu_count())
    shape, d = shape_of(dataset)
    B = 100
    torch.manual_seed(20211212)
    g = torch.Generator().manual_seed(1)
    knobs = knobs or {}
    if algo in ("flgan", "fegan"):
        from oracle.rounds import OracleFeGAN, OracleFL
        W = sample_clients
        real = torch.tanh(torch.randn(W, B, d, generator=g))
        n_real = torch.full((W,), B, dtype=torch.int32)
        if algo == "flgan":
            orc = OracleFL(W, B, shape)
            orc.load_global()
        else:
            orc = OracleFeGAN(W, B, shape, [0.1] * W, [list(range(W))])
        times = []
        for r in range(warmup + rounds):
            z_d, z_g = torch.randn(W, B, 100, generator=g), torch.randn(W, B, 100, generator=g)
            t0 = time.perf_counter()
            if algo == "flgan":
                orc.local_minibatch(real, n_real, z_d, z_g)
                orc.aggregate()
            else:
                orc.round([(real, n_real, z_d, z_g)])
            if r >= warmup:
                times.append(time.perf_counter() - t0)
        total = sum(times)
        return W * len(times) / total, total / len(times)
    from oracle.rounds import OracleMD
    W, S = sample_clients, max(1, sample_clients // per)
    orc = OracleMD(algo, W, S, B, shape, iid=1, weights_init=(algo == "mixed"), **knobs)
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


def cpu_baseline_record(algo, dataset, per, sample, rounds, warmup, knobs=None):
    rate, sec = cpu_round_rate(algo, dataset, per, sample, rounds, warmup, knobs)
    return {"value": rate, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": f"{sample} clients / {max(1, sample // per)} servers of the same round, {rounds} rounds after {warmup} "
                      f"warm-up ({sec * 1e3:.1f} ms/round); torch CPU fp32, all host threads; per-client cost is flat, so "
                      f"the rate carries over to the full client count"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", args.gpus))
    sample = args.cpu_sample_clients
    rate, sec = cpu_round_rate(args.algo, args.dataset, args.clients_per_server, sample, args.steps, args.warmup)
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
        self.nvml, self._stop = None, threading.Event()

    def _poll_nvml(self, pynvml, handle):
        # NVML in-process: a sample every 5 ms (nvidia-smi -lms 100 yields one to three samples in a 0.3 s timed region)
        bits = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
        while not self._stop.is_set():
            try:
                sm = pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self.nvml["sm"].append(float(sm))
                for n, b in bits:
                    if r & b:
                        self.nvml["reasons"].add(n)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml = {"sm": [], "reasons": set(), "max": float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM))}
            self.thread = threading.Thread(target=self._poll_nvml, args=(pynvml, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
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
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=2)
            sm = sorted(self.nvml["sm"])
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.nvml["max"],
                    "reasons": sorted(self.nvml["reasons"]), "samples": len(sm), "source": "nvml, 5 ms"}
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
class Harness:
    """Device, communicator and the timing loop shared by the headline workload and the sub-records."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        import __graft_entry__ as ge
        self.torch, self.dist, self.args = torch, dist, args
        self.rank = int(os.environ.get("RANK", 0))
        self.local_rank = int(os.environ.get("LOCAL_RANK", 0))
        self.world = int(os.environ.get("WORLD_SIZE", 1))
        if self.rank == 0:
            ge.build()
        torch.cuda.set_device(self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            dist.barrier()
        if self.rank != 0:
            ge.build()
        from cgl_gan_b200 import abi
        from cgl_gan_b200.dist import ShardComm
        self.abi = abi
        abi.require_device()
        self.dev = torch.device("cuda", self.local_rank)
        self.comm = ShardComm() if self.world > 1 else None
        peaks = {}
        p = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(p):
            peaks = json.load(open(p))
        self.hbm_peak = peaks.get("hbm_gbs", 6650.0)
        self.tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        self.peak_src = ("measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks
                         else "fallback (B200_PROFILING.md: 6650 GB/s, 1.4 PF sustained)")
        self.traffic = {}
        for name in ("traffic_r2.json", "traffic_r1.json"):     # dram bytes per launch, from the ncu capture of a round
            tpath = os.path.join(ROOT, "profiles", name)
            if os.path.exists(tpath):
                self.traffic, self.traffic_src = json.load(open(tpath)), "profiles/" + name
                break

    def sync_all(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup, profile=False):
        """W untimed + exactly K timed calls of fn(i), bracketed by barrier + synchronize, CUDA events on the current
        stream, max over ranks. -> (ms for the K steps, engine launches inside the timed region)."""
        torch, abi = self.torch, self.abi
        for i in range(warmup):
            fn(i)
        self.sync_all()
        if profile:                        # (re)start the per-kernel event pairs: the timed steps only
            abi.profile_enable(True)
        l0 = abi.launch_count()
        start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        for i in range(steps):
            fn(warmup + i)
        stop.record()
        self.sync_all()
        launches = abi.launch_count() - l0
        ms = torch.tensor([start.elapsed_time(stop)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return ms.item(), launches

    def kernel_table(self, steps):
        """cgl_profile_summary -> per kernel class: ms / launches per step, algorithmic GB/s and TFLOP/s."""
        kernels = self.abi.profile_summary()
        self.abi.profile_enable(False)
        table = {}
        for name, k in kernels.items():
            sec = k["ms"] / 1e3
            table[name] = {"ms_per_round": k["ms"] / steps, "launches_per_round": k["launches"] / steps,
                           "ms_per_launch": k["ms"] / k["launches"],
                           "algorithmic_GBps": (k["bytes"] / sec / 1e9) if sec > 0 else None,
                           "algorithmic_TFLOPps": (k["flops"] / sec / 1e12) if sec > 0 else None,
                           "bytes_per_launch": k["bytes"] / k["launches"], "flops_per_launch": k["flops"] / k["launches"]}
        return table

    def roofline_of(self, table, name, ms_per_step, dataset):
        dk = table[name]
        hbm_bound = ("[tcgen05]" not in name and "[ffma]" not in name) or "wgrad+adam" in name
        if hbm_bound:
            gbps = dk["algorithmic_GBps"] or 0.0
            r = {"kernel": name, "bound": "hbm", "achieved": gbps, "peak": self.hbm_peak, "unit": "GB/s",
                 "frac": gbps / self.hbm_peak}
        elif "[tcgen05]" in name:
            # 3xTF32: every fp32 multiply-add is three tf32 MMAs, and tf32 runs at half the bf16 rate, so the tensor
            # pipe does 6 bf16-rate units of work per algorithmic fp32 FLOP; the peak is the measured dense bf16 rate
            issued = 6.0 * dk["algorithmic_TFLOPps"]
            r = {"kernel": name, "bound": "tensor", "achieved": issued, "peak": self.tf_peak, "unit": "TFLOP/s",
                 "frac": issued / self.tf_peak, "fp32_equivalent_TFLOPps": dk["algorithmic_TFLOPps"],
                 "hbm_frac_on_its_own_bytes": (dk["algorithmic_GBps"] or 0.0) / self.hbm_peak,
                 "note": "achieved = 6 x fp32-equivalent FLOP/s: 3 tf32 MMAs per product at half the bf16 rate (the work the "
                         "tensor pipe executes for strict-fp32 results), against the measured dense bf16 peak"}
        else:
            r = {"kernel": name, "bound": "fp32 FFMA", "achieved": dk["algorithmic_TFLOPps"], "peak": 75.0, "unit": "TFLOP/s",
                 "frac": (dk["algorithmic_TFLOPps"] or 0.0) / 75.0,
                 "note": "exact-fp32 FFMA kernel; peak = 148 SMs x 128 FMA/clk x 1.965 GHz"}
        tr = self.traffic.get(dataset, {}).get(name, {}).get("dram_bytes_per_launch")
        r.update({"traffic": tr,
                  "traffic_source": (self.traffic_src + " (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch of "
                                     "this class, headline workload)") if tr is not None else None,
                  "peak_source": self.peak_src, "ms_per_launch": dk["ms_per_launch"],
                  "share_of_round": dk["ms_per_round"] / ms_per_step,
                  "bytes_per_launch": dk["bytes_per_launch"], "flops_per_launch": dk["flops_per_launch"],
                  "timing": "cudaEvent pairs on the launching stream around every launch of this kernel class, "
                            "over the timed steps of this run"})
        return r

    def rooflines(self, table, ms_per_step, dataset):
        """The kernel class with the LARGEST share of the step first (no tie-break), the runner-up second."""
        order = sorted(table, key=lambda n: -table[n]["ms_per_round"])
        first = self.roofline_of(table, order[0], ms_per_step, dataset) if order else None
        second = self.roofline_of(table, order[1], ms_per_step, dataset) if len(order) > 1 else None
        return first, second

    def free(self):
        gc.collect()
        self.torch.cuda.empty_cache()


def build_md_sim(h, algo, dataset, C, per, knob_kw=None, total_clients=None):
    """An MDStyleSim of C clients / C // per servers on this rank, random-init reference architectures tiled from a few
    distinct modules; the communicator is attached when the job has several ranks."""
    torch = h.torch
    from cgl_gan_b200 import abi, models
    from cgl_gan_b200.sim import Knobs, MDStyleSim
    shape, d = shape_of(dataset)
    S = C // per
    kw = dict(num_workers=C, num_servers=S, batch_size=100, epoch=1, cloud_epoch=1, segema=0.0, iid=1, img_shape=shape)
    kw.update(knob_kw or {})
    k = Knobs(**kw)
    tot = C * h.world if total_clients is None else total_clients
    sim = MDStyleSim(algo, k, part_sizes=[3000] * C, device=h.dev, comm=h.comm, server_offset=h.rank * S,
                     total_data_len=3000 * tot)
    g_proto = [sim.G.make_module() for _ in range(4)]
    if algo == "mixed":
        from cgl_gan_b200.models import weights_init
        for m in g_proto:
            m.apply(weights_init)
    d_arch = abi.ARCH_D_2D if d == 2 else (abi.ARCH_D_MNIST2 if sim.loss_kind == abi.LOSS_CE else abi.ARCH_D_MNIST1)
    d_proto = [models.Discriminator(shape, arch=d_arch) for _ in range(8)]
    sim.G.load_modules([g_proto[s % 4] for s in range(S)])
    sim.bank.load_modules([d_proto[c % 8] for c in range(C)])
    return sim


def synthetic_batches(h, C, d, ring=2, seed=1234):
    torch = h.torch
    gen = torch.Generator().manual_seed(seed + h.rank)
    host = [torch.tanh(torch.randn(C, 100, d, generator=gen)).pin_memory() for _ in range(ring)]
    return host, [x.to(h.dev) for x in host]


def measure_md(h, algo, dataset, C, per, steps, warmup, knob_kw=None, total_clients=None, graph=False):
    """Device-resident timing of an MD-style round + its kernel table."""
    torch = h.torch
    shape, d = shape_of(dataset)
    sim = build_md_sim(h, algo, dataset, C, per, knob_kw, total_clients)
    _, dev_real = synthetic_batches(h, C, d)
    n_real = torch.full((C,), 100, dtype=torch.int32, device=h.dev)
    fn = (lambda i: sim.round_graph(dev_real[i % 2], n_real)) if graph else (lambda i: sim.round(dev_real[i % 2], n_real))
    ms, launches = h.timed(fn, steps, warmup, profile=not graph)
    table = h.kernel_table(steps) if not graph else {}
    del sim, dev_real
    h.free()
    return ms, launches, table


def sub_record(h, name, workload, clients_total, ms, steps, launches, table, dataset, scaling, extra=None):
    ms_step = ms / steps
    rec = {"workload": workload, "value": clients_total * steps / (ms / 1e3), "unit": UNIT, "ms_per_step": ms_step,
           "n_gpus": h.world, "scaling": scaling, "steps": steps, "warmup": SUB_WARMUP, "gpu_launches_per_step": launches / steps}
    if table:
        first, second = h.rooflines(table, ms_step, dataset)
        rec["roofline"], rec["roofline_second_kernel"] = first, second
        rec["kernel_ms_sum_per_step"] = sum(v["ms_per_round"] for v in table.values())
    if extra:
        rec.update(extra)
    return rec


def run_sub_records(h, args, wanted):
    """The other BASELINE.json workloads, measured in the same run. Every rank takes part in every record that has a
    collective; rank 0 reports."""
    torch = h.torch
    out = {}
    want = (lambda n: wanted == "all" or n in wanted.split(","))
    cpu = (not args.no_cpu_baseline) and h.world == 1 and h.rank == 0
    W = h.world

    def guarded(name, fn):
        if not want(name):
            return
        try:
            out[name] = fn()
        except Exception as e:            # a failing sub-record must not take the headline line down with it
            out[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
            h.free()

    # ---- strong scaling: 1024 clients IN TOTAL over the ranks (north_star) ------------------------------------
    def strong(dataset, per):
        C = 1024 // W
        ms, launches, table = measure_md(h, "cglgan", dataset, C, per, SUB_STEPS, SUB_WARMUP, total_clients=1024)
        extra = {"clients_per_gpu": C}
        if W > 1:
            # what limits the small per-GPU round: the same round without the cloud exchange (no all-reduce) ...
            ms_nc, _, _ = measure_md(h, "cglgan", dataset, C, per, SUB_STEPS, SUB_WARMUP, knob_kw={"cloud_epoch": 0},
                                     total_clients=1024)
            extra["ms_per_step_without_cloud_exchange"] = ms_nc / SUB_STEPS
            extra["cloud_exchange_ms"] = (ms - ms_nc) / SUB_STEPS
        if dataset == "2dmg":
            # K1: the same rounds with the OTHER client-step implementation (the one that is not the default): the fused
            # shared-memory-resident kernel of csrc/client_fused.cuh (one launch per client step) against the layered kernels
            was = h.abi.lib.cgl_get_fused_client_step()
            try:
                h.abi.check(h.abi.lib.cgl_set_fused_client_step(0 if was else 1))
                ms_k, launches_k, table_k = measure_md(h, "cglgan", dataset, C, per, SUB_STEPS, SUB_WARMUP, total_clients=1024)
                key = "layered_client_step" if was else "fused_client_step"
                extra[key] = {"ms_per_step": ms_k / SUB_STEPS, "value": 1024 * SUB_STEPS / (ms_k / 1e3),
                              "gpu_launches_per_step": launches_k / SUB_STEPS,
                              "kernels_ms_per_step": {k: v["ms_per_round"] for k, v in table_k.items()}}
                extra["client_step_default"] = "fused (csrc/client_fused.cuh)" if was else "layered (grouped GEMM kernels)"
            finally:
                h.abi.check(h.abi.lib.cgl_set_fused_client_step(was))
        # ... and the same rounds replayed from a CUDA graph (the cloud all-reduce captured with them): no launch gaps
        try:
            ms_g, _, _ = measure_md(h, "cglgan", dataset, C, per, SUB_STEPS, SUB_WARMUP + 2, total_clients=1024, graph=True)
            extra["ms_per_step_cuda_graph"] = ms_g / SUB_STEPS
            extra["value_cuda_graph"] = 1024 * SUB_STEPS / (ms_g / 1e3)
        except Exception as e:
            extra["cuda_graph_error"] = f"{type(e).__name__}: {e}"[:200]
            h.free()
        rec = sub_record(h, "strong", f"CGLGAN {dataset.upper()} round, 1024 clients / {1024 // per} servers IN TOTAL dealt over "
                         f"{W} GPU(s), batch 100, cloud all-reduce every round; value = eager rounds (per-kernel events), "
                         f"value_cuda_graph = MDStyleSim.round_graph", 1024, ms, SUB_STEPS, launches, table, dataset,
                         "strong", extra)
        if "kernel_ms_sum_per_step" in rec:
            # ... and the host side: time of the round not covered by engine kernels (launch gaps, ATen weight math)
            rec["ms_outside_engine_kernels"] = rec["ms_per_step"] - rec["kernel_ms_sum_per_step"]
        if cpu:
            rec["cpu_baseline"] = cpu_baseline_record("cglgan", dataset, per, 16, 5, 2)
        return rec
    guarded("strong_1024_mnist", lambda: strong("mnist", 4))
    guarded("strong_1024_2dmg", lambda: strong("2dmg", 2))

    # ---- configs[0]: CGLGAN 2DMG, repo-default topology, replayed from a CUDA graph -----------------------------------------
    def repo_default():
        keep, h.comm = h.comm, None              # 10 workers do not shard: every rank runs a replica, rank 0 reports
        try:
            ms_e, launches, table = measure_md(h, "cglgan", "2dmg", 10, 2, 20, 5)
            ms_g, _, _ = measure_md(h, "cglgan", "2dmg", 10, 2, 20, 5, graph=True)
        finally:
            h.comm = keep
        rec = sub_record(h, "2dmg", "CGLGAN 2DMG round, repo-default 10 workers / 5 servers, batch 100 (BASELINE configs[0]); "
                         "value = CUDA-graph replay (MDStyleSim.round_graph), replicas only", 10, ms_g, 20, launches, table,
                         "2dmg", "replicas", {"ms_per_step_eager": ms_e / 20, "value_eager": 10 * 20 / (ms_e / 1e3)})
        rec["n_gpus"] = 1
        if cpu:
            rec["cpu_baseline"] = cpu_baseline_record("cglgan", "2dmg", 2, 10, 20, 3)
        return rec
    guarded("2dmg_repo_default", repo_default)

    # ---- configs[2]: CAPGAN + Mix-G with the neighbour discriminator share every E = 5 rounds --------------------------------
    def mixg():
        C = args.clients
        kw = {"E": 5, "d_share": "group_mean", "segema": 0.5, "num_communication": 20000}
        ms, launches, table = measure_md(h, "mixed", "mnist", C, 4, 10, 5, knob_kw=kw)
        rec = sub_record(h, "mixg", f"mixed-gan.py round (CAPGAN weighting + MixGenerator, weights_init, CE x0.5, segema 0.5), {C} "
                         f"clients / {C // 4} servers per GPU, neighbour discriminator share (group mean) every E = 5 rounds: 2 of "
                         f"the 10 timed rounds share", C * W, ms, 10, launches, table, "mnist", "weak")
        rec["warmup"] = 5
        if cpu:
            rec["cpu_baseline"] = cpu_baseline_record("mixed", "mnist", 4, 16, 5, 2, {"E": 5, "segema": 0.5})
        return rec
    guarded("capgan_mixg_E5", mixg)

    # ---- configs[3]: MD-GAN, ONE server, clients dealt over the GPUs ------------------------------------------------------------
    def mdgan():
        from cgl_gan_b200 import models
        from cgl_gan_b200.sim import Knobs, MDSingleServerSim
        shape, d = shape_of("mnist")
        Wk = 1024
        k = Knobs(num_workers=Wk, num_servers=1, batch_size=100, img_shape=shape)
        sim = MDSingleServerSim("mdgan", k, part_sizes=[3000] * Wk, device=h.dev, comm=h.comm, rank=h.rank, world=W)
        C = sim.hi - sim.lo
        g_mod = sim.G.make_module()
        d_proto = [models.Discriminator(shape) for _ in range(8)]
        sim.G.load_modules([g_mod])
        sim.bank.load_modules([d_proto[c % 8] for c in range(C)])
        _, dev_real = synthetic_batches(h, C, d)
        n_real = torch.full((C,), 100, dtype=torch.int32, device=h.dev)
        ms, launches = h.timed(lambda i: sim.round(dev_real[i % 2], n_real), SUB_STEPS, SUB_WARMUP, profile=True)
        table = h.kernel_table(SUB_STEPS)
        del sim, dev_real
        h.free()
        rec = sub_record(h, "mdgan", f"MD-GAN MNIST round (MDGAN/MNIST/mdgan.py), ONE server, 1024 clients dealt over {W} GPU(s): "
                         "replicated generator, all-gather of the losses, all-reduce of sum_i w_i dLoss_i/dXg ([100, 784])",
                         1024, ms, SUB_STEPS, launches, table, "mnist", "strong", {"clients_per_gpu": C})
        if cpu:
            rec["cpu_baseline"] = cpu_baseline_record("mdgan", "mnist", 16, 16, 5, 2)
        return rec
    guarded("mdgan_single_server", mdgan)

    # ---- configs[3]: FL-GAN ---------------------------------------------------------------------------------------------------------
    def flgan():
        from cgl_gan_b200 import models
        from cgl_gan_b200.sim import FLStyleSim, Knobs
        shape, d = shape_of("mnist")
        C = args.clients
        sim = FLStyleSim(Knobs(num_workers=C, num_servers=1, batch_size=100, img_shape=shape), device=h.dev, comm=h.comm)
        g_proto = [sim.G.make_module() for _ in range(4)]
        d_proto = [models.Discriminator(shape) for _ in range(8)]
        sim.G.load_modules([g_proto[c % 4] for c in range(C)])
        sim.bank.load_modules([d_proto[c % 8] for c in range(C)])
        _, dev_real = synthetic_batches(h, C, d)
        n_real = torch.full((C,), 100, dtype=torch.int32, device=h.dev)

        def step(i):
            sim.local_minibatch(dev_real[i % 2], n_real)
            sim.aggregate()
        ms, launches = h.timed(step, SUB_STEPS, SUB_WARMUP, profile=True)
        table = h.kernel_table(SUB_STEPS)
        del sim, dev_real
        h.free()
        rec = sub_record(h, "flgan", f"FL-GAN MNIST step: {C} clients/GPU, one local minibatch (cgl_fl_step: D step + G step, batch "
                         "100) on every client, then the FedAvg of every G and D (FLGAN/MNIST/flgan.py:143-163,245-270)",
                         C * W, ms, SUB_STEPS, launches, table, "mnist", "weak")
        if cpu:
            rec["cpu_baseline"] = cpu_baseline_record("flgan", "mnist", 1, 8, 4, 1)
        return rec
    guarded("flgan", flgan)

    # ---- configs[3]: FeGAN, frac_workers = 0.2 -------------------------------------------------------------------------------------
    def fegan():
        import numpy as np
        from cgl_gan_b200 import models
        from cgl_gan_b200.partition import init_groups
        from cgl_gan_b200.sim import FeGANSim, Knobs
        shape, d = shape_of("mnist")
        P = 1024
        rs = np.random.RandomState(7)
        freq = [np.bincount(rs.randint(0, 10, size=3), minlength=10) * 100 for _ in range(P)]   # ~3 classes per client
        groups, _ = init_groups(P, freq, 0.2, max_groups=64)
        sk = (rs.rand(P) * 0.5).tolist()
        sim = FeGANSim(Knobs(num_workers=P, num_servers=1, batch_size=100, img_shape=shape), sk, groups, device=h.dev,
                       comm=h.comm, rank=h.rank, world=W)
        C = sim.hi - sim.lo
        g_proto = [sim.G.make_module() for _ in range(4)]
        d_proto = [models.Discriminator(shape) for _ in range(8)]
        sim.G.load_modules([g_proto[c % 4] for c in range(C)])
        sim.bank.load_modules([d_proto[c % 8] for c in range(C)])
        sim.load_global(g_proto[0], d_proto[0])
        gsz = len(groups[0])
        gen = torch.Generator().manual_seed(11)
        real_all = torch.tanh(torch.randn(gsz, 100, d, generator=gen)).to(h.dev)
        served = []

        def step(i):
            mine, ids = sim.begin_round()
            if mine:
                sim.local_minibatch(real_all[:len(mine)], None, client_ids=ids)
            sim.end_round(mine, ids)
            served.append(len(mine))
        ms, launches = h.timed(step, SUB_STEPS, SUB_WARMUP, profile=True)
        table = h.kernel_table(SUB_STEPS)
        del sim, real_all
        h.free()
        rec = sub_record(h, "fegan", f"FeGAN MNIST round (fegan.py:125-165): population 1024 dealt over {W} GPU(s), frac_workers 0.2 "
                         f"-> init_groups picks {gsz} clients per round; they load the global G / D, run one local minibatch, and "
                         "the softmax(sk)-weighted FedAvg replaces the global vectors", gsz, ms, SUB_STEPS, launches, table,
                         "mnist", "strong", {"group_size": gsz, "served_by_rank0_last_round": served[-1]})
        if cpu:
            rec["cpu_baseline"] = cpu_baseline_record("fegan", "mnist", 1, 8, 4, 1)
        return rec
    guarded("fegan", fegan)
    return out


def run_b200(args):
    h = Harness(args)
    torch, dist, abi = h.torch, h.dist, h.abi
    rank, world, dev = h.rank, h.world, h.dev
    shape, d = shape_of(args.dataset)
    C, per, B = args.clients, args.clients_per_server, 100
    S = C // per
    torch.manual_seed(20211212 + rank)
    fl = args.algo == "flgan"
    if fl:
        from cgl_gan_b200 import models
        from cgl_gan_b200.sim import FLStyleSim, Knobs
        sim = FLStyleSim(Knobs(num_workers=C, num_servers=1, batch_size=B, img_shape=shape), device=dev, comm=h.comm)
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
        sim.profile = False
    else:
        sim = build_md_sim(h, args.algo, args.dataset, C, per)

    # synthetic MNIST-shaped batches: a ring of pinned host buffers (e2e) and device-resident copies (value)
    ring = 2
    host, dev_real = synthetic_batches(h, C, d, ring)
    n_real_dev = torch.full((C,), B, dtype=torch.int32, device=dev)
    stage = [torch.empty(C, B, d, device=dev) for _ in range(2)]
    loss_host = (torch.empty(C, dtype=torch.float32) if fl else torch.empty(S, per, dtype=torch.float32)).pin_memory()
    copy_stream = torch.cuda.Stream()
    pending = [None]

    def resident_round(i):
        return sim.round(dev_real[i % ring], n_real_dev)

    def e2e_prime():
        with torch.cuda.stream(copy_stream):
            stage[0].copy_(host[0], non_blocking=True)
            ev = torch.cuda.Event(); ev.record(copy_stream)
        pending[0] = (stage[0], ev)

    def e2e_round(i):
        """H2D of this round's real batches from pinned memory (prefetched one round ahead on a copy stream),
        the round, D2H of the clients' G losses."""
        cur = torch.cuda.current_stream()
        buf, ev = pending[0]
        cur.wait_event(ev)
        nxt = stage[(i + 1) % 2]
        with torch.cuda.stream(copy_stream):   # stage[(i+1)%2] is free: round i-1 was synchronised
            nxt.copy_(host[(i + 1) % ring], non_blocking=True)
            ev2 = torch.cuda.Event()
            ev2.record(copy_stream)
        loss = sim.round(buf, n_real_dev)
        loss_host.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()     # the user reads the round's losses
        pending[0] = (nxt, ev2)

    sampler = ClockSampler(h.local_rank)
    if rank == 0:
        sampler.start()
    # CUDA-event pairs around every engine kernel, on the launching stream, over the timed rounds
    if not fl:
        sim.profile = True
    ms, launches = h.timed(resident_round, args.steps, args.warmup, profile=True)
    client_ms = sim.client_step_ms() if not fl else None
    sim.profile = False
    table = h.kernel_table(args.steps)
    clocks = sampler.stop() if rank == 0 else None
    e2e_prime()
    ms_e2e, _ = h.timed(e2e_round, args.steps, args.warmup)

    # e2e through the path a user runs: the dataset resident in HBM, DataLoader-exact row ids (8 B / sample) from the
    # host every round, cgl_gather_rows on the device, the round, D2H of the losses
    e2e_res = None
    if not fl:
        from cgl_gan_b200.data import ResidentPartitions
        gen = torch.Generator().manual_seed(77 + rank)
        n_data = 60000
        data = torch.tanh(torch.randn(n_data, d, generator=gen))
        perm = torch.randperm(n_data, generator=gen)
        parts = [perm[(c * 3000) % n_data:(c * 3000) % n_data + 3000] for c in range(C)]
        rp = ResidentPartitions(data, parts, B, device=dev, shuffle=True)

        def resident_e2e_round(i):
            real, n_real = rp.next_batches()
            loss = sim.round(real, n_real)
            loss_host.copy_(loss, non_blocking=True)
            rp.prefetch()                                  # the next round's row ids are drawn while this round runs
            torch.cuda.current_stream().synchronize()
        ms_res, _ = h.timed(resident_e2e_round, args.steps, args.warmup)
        e2e_res = {"value": C * world * args.steps / (ms_res / 1e3), "unit": UNIT, "ms_per_step": ms_res / args.steps,
                   "h2d_bytes_per_step": int(rp.h2d_bytes_per_round * world), "d2h_bytes_per_step": int(S * per * 4 * world),
                   "path": "data.ResidentPartitions: dataset resident in HBM, DataLoader(shuffle=True)-exact row ids drawn on the "
                           "host every round, cgl_gather_rows, MDStyleSim.round, D2H of the losses"}
        del rp, data

    total_clients = C * world
    value = total_clients * args.steps / (ms / 1e3)
    e2e_value = total_clients * args.steps / (ms_e2e / 1e3)
    ms_step = ms / args.steps
    roofline, roofline_2 = h.rooflines(table, ms_step, args.dataset)
    client = None
    if client_ms and d == 784:
        client = {"client_step_ms_per_round": client_ms, "algorithmic_bytes_per_client_step": BYTES_PER_CLIENT_STEP,
                  "hbm_GBps": BYTES_PER_CLIENT_STEP * C / (client_ms / 1e3) / 1e9,
                  "hbm_frac": BYTES_PER_CLIENT_STEP * C / (client_ms / 1e3) / 1e9 / h.hbm_peak,
                  "fp32_TFLOPps": FLOPS_PER_CLIENT_STEP * C / (client_ms / 1e3) / 1e12}
    del sim, dev_real, host, stage
    h.free()

    configs = {}
    if args.configs != "none":
        configs = run_sub_records(h, args, args.configs)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "round_time_s": ms_step / 1e3,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(C * B * d * 4 * world), "d2h_bytes_per_step": int(loss_host.numel() * 4 * world),
                    "path": "raw batches: H2D of every client's [100, d] real batch from pinned memory each round"},
            "e2e_resident": e2e_res,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "roofline_second_kernel": roofline_2,
            "client_step": client,
            "kernels": table,
            "configs": configs,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_record(args.algo, args.dataset, per, args.cpu_sample_clients, 20, 3)
        print(json.dumps(line))
    if h.comm is not None:
        h.comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
