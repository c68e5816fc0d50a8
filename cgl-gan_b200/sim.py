"""Host loops: the reference's simulations as single-threaded rounds over the engine.

The reference runs one Python thread per Worker / Server / Cloud and moves tensors through queues
(SURVEY.md 3.2-3.3). Here one `round()` advances every server and every client of this process:

  MD-style (CGLGAN, CAPGAN, Mix-G, MDGAN, ACGAN): MDStyleSim
     Server.train  CGLGAN/2DMG/main.py:225-278, capgan.py:211-262, mixed-gan.py:238-292,
                   MDGAN/MNIST/mdgan.py:180-207, ACGAN/MNIST/acgan.py:149-179
     Worker.train  CGLGAN/2DMG/main.py:344-375, capgan.py:316-349
     Cloud.run     CGLGAN/2DMG/main.py:116-136 (+ the segema mix at :205-208)
  FL-style (FLGAN): FLStyleSim
     Worker.train  FLGAN/MNIST/flgan.py:245-270, FLGAN/2DMG/flgan.py:227-258
     Server.run    FLGAN/MNIST/flgan.py:143-163

Knob names are the reference's module-level globals (README.md:23-35).
"""
import ctypes
from dataclasses import dataclass, field, replace
from random import Random
from typing import Optional

import torch
import torch.nn.functional as F

from . import abi
from .engine import ClientBank, _stream
from .generators import StackedGenerator
from .partition import assign_clients

ALGOS = {
    #  name        : (multi_head, d_arch(2d, mnist),                   loss,          d_scale, weighting)
    "cglgan": dict(loss=abi.LOSS_BCE, d_scale=1.0, weighting="cgl"),
    "capgan": dict(loss=abi.LOSS_CE, d_scale=0.5, weighting="cap"),
    "capgan_copy": dict(loss=abi.LOSS_CE, d_scale=0.5, weighting="cap_copy"),
    "mixed": dict(loss=abi.LOSS_CE, d_scale=0.5, weighting="mixed"),
    "mdgan": dict(loss=abi.LOSS_BCE, d_scale=1.0, weighting="mean"),
    "acgan": dict(loss=abi.LOSS_CE, d_scale=1.0, weighting="mean"),
}


@dataclass
class Knobs:
    """The reference's global knobs (CGLGAN/2DMG/main.py:30-58, mixed-gan.py:41-59)."""
    num_communication: int = 20000   # the round counter t counts DOWN from here (while t > 0: ...; t -= 1)
    num_workers: int = 10
    num_servers: int = 5
    batch_size: int = 100
    epoch: int = 1            # local D steps per round
    cloud_epoch: int = 1      # rounds between cloud aggregations (0: never); root capgan.py: EPOCHS over the server's data
    segema: float = 0.0       # 1: fully independent servers, 0: fully shared trunk
    E: int = 0                # rounds between neighbour-D shares (0: off, as shipped: commented out)
    frac_workers: float = 1.0
    iid: int = 1
    b1: float = 0.5
    b2: float = 0.999
    lr_g: float = 0.0002
    lr_d: float = 0.0002
    img_shape: tuple = (2,)
    cloud_mode: str = "intended"   # "as_written": the state_dict key mismatch makes Cloud a no-op
    d_share: str = "group_mean"    # neighbour-D share flavour when E > 0: "group_mean" | "swap"


class MDStyleSim:
    """Server owns G, every client owns a D. One instance per process (= per GPU shard)."""

    def __init__(self, algo, knobs: Knobs, part_sizes=None, device="cuda", comm=None, server_offset=0,
                 total_data_len=None):
        assert algo in ALGOS
        self.algo, self.k = algo, knobs
        spec = ALGOS[algo]
        k = knobs
        self.device = torch.device(device)
        self.S = k.num_servers
        self.N = k.num_workers // k.num_servers
        self.C = self.S * self.N
        self.B = k.batch_size
        d = 1
        for s in k.img_shape:
            d *= s
        self.d = d
        two_d = d == 2
        # generator flavour: CGLGAN -> Generator(ims, N if iid != 0 else 1) (CGLGAN/2DMG/main.py:191);
        # mixed-gan -> MixGenerator(ims, N); the rest -> plain Generator(ims)
        if algo == "cglgan":
            self.n_heads = self.N if k.iid != 0 else 1
        elif algo == "mixed":
            self.n_heads = self.N
        else:
            self.n_heads = 0
        self.multi_head = self.n_heads == self.N and self.n_heads > 0 and not (algo == "cglgan" and k.iid == 0)
        if two_d:
            d_arch = abi.ARCH_D_2D
            loss = abi.LOSS_BCE            # every 2DMG script pairs BCE with the sigmoid D
        else:
            loss = spec["loss"]
            d_arch = abi.ARCH_D_MNIST2 if loss == abi.LOSS_CE else abi.ARCH_D_MNIST1
        self.loss_kind = loss
        self.G = StackedGenerator(k.img_shape, self.S, self.n_heads, device=self.device, lr=k.lr_g, b1=k.b1, b2=k.b2)
        self.bank = ClientBank(d_arch, self.C, self.B, device=self.device, loss_kind=loss,
                               d_loss_scale=1.0 if two_d else spec["d_scale"], lr=k.lr_d, b1=k.b1, b2=k.b2)
        self.weighting = spec["weighting"]
        self.client_list, _ = assign_clients(self.C, self.S)
        self.server_of = torch.arange(self.C, device=self.device, dtype=torch.int32) // self.N
        self.srv_ptr = torch.arange(0, self.C + 1, self.N, device=self.device, dtype=torch.int32)
        # beta: client share of its server's data; A: server share of all data (CGLGAN/2DMG/main.py:184-188,117-122)
        if part_sizes is None:
            part_sizes = [1] * self.C
        sizes = torch.tensor(part_sizes, dtype=torch.float32).view(self.S, self.N)
        self.data_len = sizes.sum(1)
        self.beta = (sizes / self.data_len.unsqueeze(1)).to(self.device)
        tot = self.data_len.sum() if total_data_len is None else torch.tensor(float(total_data_len))
        self.A = (self.data_len / tot).to(self.device)
        self.Lambda = torch.zeros(self.S, device=self.device)
        self.comm = comm                  # dist.ShardComm or None (single process)
        self.server_offset = server_offset
        self.t = 0                        # rounds done
        self.last_F_max = None
        self.swap_rd = [Random(s + server_offset + 100) for s in range(self.S)]  # Server.rd, main.py:154-155
        self.profile = False              # bench: CUDA events around the client-step kernels
        self._events = []
        # overlap_g: the generators' Xg pass runs on a second stream under the clients' D step (it is only needed by the G loss
        # that follows the D step). Same kernels, same inputs, same order on every buffer: bit-identical results.
        # (env CGL_OVERLAP_G=1 switches it on for measurements; measured: profiles/overlap_g_r2.log)
        import os
        self.overlap_g = os.environ.get("CGL_OVERLAP_G", "0") == "1"
        self._side = None

    def client_step_ms(self):
        """Mean device time per round of the client-step calls (cgl_d_step + cgl_g_loss), from the CUDA
        events recorded while self.profile was set. Call after a synchronize."""
        if not self._events:
            return None
        ms = sum(a.elapsed_time(b) for a, b in self._events) / len(self._events)
        self._events = []
        return ms

    # ---- initialisation from reference-style modules -------------------------------------------
    def load(self, g_modules, d_modules):
        self.G.load_modules(g_modules)
        self.bank.load_modules(d_modules)

    # ---- one communication round ----------------------------------------------------------------
    def round(self, real, n_real=None, z_d=None, z_g=None):
        """real: [epoch, C, B, d] (or [C, B, d] when epoch == 1) device tensor of the clients' minibatches,
        n_real: matching valid-row counts (None: full batches). Returns the clients' G losses [S, N]."""
        k, S, N, B, d = self.k, self.S, self.N, self.B, self.d
        if real.dim() == 3:
            real = real.unsqueeze(0)
        if n_real is not None and n_real.dim() == 1:
            n_real = n_real.unsqueeze(0)
        if self._cloud_due():
            self.cloud_aggregate()
        if self._share_due():
            self.share_discriminators()
        if z_d is None:
            z_d = torch.randn(S, B, 100, device=self.device)
        if z_g is None:
            z_g = torch.randn(S, B, 100, device=self.device)
        G = self.G
        Xd = G(z_d)          # no_grad pass of the reference: only its BatchNorm running statistics survive
        overlap = self.overlap_g and not self.profile and real.shape[0] == 1
        if overlap:
            # the Xg pass starts after the Xd pass (BatchNorm running statistics, shared workspace) on a second stream;
            # the main stream goes on with the D step and waits for it before the G loss
            main = torch.cuda.current_stream()
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
            self._side.wait_stream(main)
            with torch.cuda.stream(self._side):
                Xg = G(z_g)
            for t in (Xg,) + tuple(x for x in G._last if x is not None):
                t.record_stream(main)      # allocated on the side stream, consumed (and freed) on the main stream
        else:
            Xg = G(z_g)          # the pass the generator is trained through
        shared = not self.multi_head
        if shared:           # one batch per server, seen by all of its clients (CGLGAN iid==0: Generator(ims, 1))
            Xd, Xg_flat = Xd.reshape(S, B, d), Xg.reshape(S, B, d)
        else:                # head i of server s feeds client s*N+i
            Xd, Xg_flat = Xd.reshape(S * N, B, d), Xg.reshape(S * N, B, d)
        idx = self.server_of if shared else None
        if self.profile:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        E = real.shape[0]
        for e in range(E - 1):
            self.last_d_loss = self.bank.d_step(real[e], Xd, n_real=None if n_real is None else n_real[e],
                                                fake_idx=idx)
        # the last D step and the G loss through the updated D: one ABI call (one launch for the 2DMG discriminator)
        if overlap:
            self.last_d_loss = self.bank.d_step(real[0], Xd, n_real=None if n_real is None else n_real[0], fake_idx=idx)
            torch.cuda.current_stream().wait_stream(self._side)
            loss, dxg = self.bank.g_loss_raw(Xg_flat, xg_idx=idx)
        else:
            self.last_d_loss, loss, dxg = self.bank.client_step(real[E - 1], Xd, Xg_flat,
                                                                n_real=None if n_real is None else n_real[E - 1], idx=idx)
        loss = loss.view(S, N)
        if self.profile:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            self._events.append((ev0, ev1))
        w = self._server_weights(loss)
        if self.multi_head:
            # heads: d(sum_i loss_i); trunk: d(sum_i w_i loss_i)  (CGLGAN/2DMG/main.py:254-269, mixed-gan.py:263-281)
            G.backward_step(dxg.view(S, N, B, d), trunk_w=w)
        else:
            # F_max.backward() through a shared Xg: every client's dLoss/dXg, weighted, summed per server
            dy = torch.empty(S, B, d, device=self.device)
            wf = w.reshape(-1).contiguous().float()
            abi.check(abi.lib.cgl_dxg_reduce(S, abi.ptr(self.srv_ptr), None, abi.ptr(wf), abi.ptr(dxg), B * d,
                                             abi.ptr(dy), _stream()))
            self.bank.launches += 1
            G.backward_step(dy.view(S, 1, B, d) if self.n_heads else dy)
        self.t += 1
        return loss

    # ---- the same round as a CUDA graph ------------------------------------------------------------
    def round_graph(self, real, n_real=None, z_d=None, z_g=None):
        """round() replayed from a CUDA graph: for small topologies (the repo-default 10 workers / 5 servers, the 2DMG
        networks) a round is ~50 launches of 10-300 us and the host's launch rate, not the GPU, sets the pace.
        The first call runs an eager round (it also sizes every workspace), the second captures, all later calls copy
        the inputs into the captured buffers and replay. Fixed control flow only: cloud_epoch in {0, 1}, E == 0, one
        (any number of ranks); z_d / z_g must be given either always or never (never: torch.randn inside the graph)."""
        k = self.k
        assert k.E == 0 and k.cloud_epoch in (0, 1) and not self.profile, "round_graph: unsupported knobs"
        assert self.algo != "capgan" or not k.cloud_epoch, "round_graph: capgan's cloud period is decided on the host"
        # (with a communicator the cloud all-reduce is captured too: ncclAllReduce on the capturing stream; every rank
        #  captures and replays the same sequence)
        st = getattr(self, "_graph_state", None)
        if st is None:
            self._graph_state = {"graph": None}
            return self.round(real, n_real, z_d, z_g)
        if st["graph"] is not None:
            given = (n_real is not None, z_d is not None, z_g is not None)
            if given != st["given"]:
                raise ValueError(f"round_graph was captured with (n_real, z_d, z_g) given = {st['given']}, now {given}: "
                                 "a captured round cannot change which inputs it reads")
        if st["graph"] is None:
            st["given"] = (n_real is not None, z_d is not None, z_g is not None)
            assert st["given"][1] == st["given"][2], "z_d and z_g must be given together"
            st["real"] = real.clone()
            st["n_real"] = None if n_real is None else n_real.clone()
            st["z_d"] = None if z_d is None else z_d.clone()
            st["z_g"] = None if z_g is None else z_g.clone()
            t0 = self.t
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph):
                st["loss"] = self.round(st["real"], st["n_real"], st["z_d"], st["z_g"])
            self.t = t0                      # capturing records the kernels, it does not run them
            st["graph"] = graph
        else:
            st["real"].copy_(real)
            if n_real is not None:
                st["n_real"].copy_(n_real)
            if z_d is not None:
                st["z_d"].copy_(z_d)
                st["z_g"].copy_(z_g)
        st["graph"].replay()
        self.t += 1
        return st["loss"]

    def _cloud_due(self):
        """The servers' `if t % ... == 0` before Server.train, with the reference's counter t = num_communication - rounds done:
        CGLGAN/2DMG/main.py:201, mixed-gan.py:193, CAPGAN/MNIST/capgan.py:169: every cloud_epoch rounds;
        capgan.py:169: `t % (self.data_len * cloud_epoch / batch_size) == 0` -- there cloud_epoch counts EPOCHS over the
        server's data and the period is a float32 tensor (int % tensor = torch.remainder). The exchange is a rendezvous of
        all servers with the Cloud (capgan.py:108-117), so servers that disagree would deadlock the reference: refused."""
        k = self.k
        if not k.cloud_epoch:
            return False
        t_ref = k.num_communication - self.t
        if self.algo != "capgan":
            return t_ref % k.cloud_epoch == 0
        period = self.data_len * k.cloud_epoch / k.batch_size          # float32 [S], as in the reference
        due = torch.remainder(torch.tensor(float(t_ref)), period) == 0
        if self.comm is not None:                                      # every rank must take the same decision
            import torch.distributed as dist
            flags = torch.tensor([float(due.all()), float(due.any())], device=self.device)
            dist.all_reduce(flags[0:1], op=dist.ReduceOp.MIN)
            dist.all_reduce(flags[1:2], op=dist.ReduceOp.MAX)
            all_due, any_due = bool(flags[0].item()), bool(flags[1].item())
        else:
            all_due, any_due = bool(due.all()), bool(due.any())
        if any_due and not all_due:
            raise ValueError("capgan: the servers' cloud periods data_len*cloud_epoch/batch_size differ "
                             f"({period.tolist()}); the reference's Cloud rendezvous (capgan.py:108-117) would deadlock")
        return all_due

    def _share_due(self):
        """Neighbour-D share every E rounds. ACGAN counts the rounds done, `(num_communication - t) % E == 0`
        (ACGAN/MNIST/acgan.py:240: round 0 shares too); MD-GAN tests the down-counter itself, `t % E == 0`
        (MDGAN/MNIST/mdgan.py:158,258); the scripts without a call site follow MD-GAN."""
        k = self.k
        if not k.E:
            return False
        if self.algo == "acgan":
            return self.t % k.E == 0
        return (k.num_communication - self.t) % k.E == 0

    def _server_weights(self, loss):
        """Per-client weight of its G loss in the server objective; also advances Lambda."""
        beta, Lam = self.beta, self.Lambda
        kind = self.weighting
        if kind == "mean":                      # MDGAN/MNIST/mdgan.py:203, ACGAN/MNIST/acgan.py:173
            self.last_F_max = loss.mean(1)
            return torch.full_like(loss, 1.0 / loss.shape[1])
        if kind == "cgl":                       # CGLGAN/2DMG/main.py:261-274
            gamma = F.softmax(Lam.unsqueeze(1) * loss, dim=1)
            F_beta = (beta * loss).sum(1)
            F_gamma = (gamma * loss).sum(1)
            self.last_F_max = (F_beta + F_gamma) / 2
            grad = (loss * loss * gamma).sum(1) - (loss * gamma * F_gamma.unsqueeze(1)).sum(1)
            new_lambda = Lam + 10 * grad
            self.Lambda.copy_(new_lambda)     # in place: a captured round (round_graph) must find its state where it left it
            return (beta + gamma) / 2
        if kind == "cap":                       # capgan.py:247-249
            alpha = F.softmax(Lam.unsqueeze(1) * loss, dim=1)
            alpha = F.softmax(alpha * beta, dim=1)
        elif kind == "cap_copy":                # CAPGAN/MNIST/capgan.py:241-243
            gamma = F.softmax(Lam.unsqueeze(1) * loss, dim=1)
            alpha = F.softmax(beta * gamma, dim=1)
        elif kind == "mixed":                   # mixed-gan.py:276-277
            alpha = F.softmax(beta * Lam.unsqueeze(1) * loss, dim=1)
        else:
            raise ValueError(kind)
        self.last_F_max = (alpha * loss).sum(1) - 0.001 * Lam
        # opti_L = SGD([Lambda], lr=0.1); dF_max/dLambda = -0.001 (capgan.py:160,250,259)
        self.Lambda.copy_(Lam.add(torch.full_like(Lam, -0.001), alpha=-0.1))
        return alpha

    # ---- aggregation ------------------------------------------------------------------------------
    def cloud_aggregate(self):
        """Cloud.run + the receiving half of Server.run (CGLGAN/2DMG/main.py:124-136,201-208):
        p = sum_s A[s] * trunk_s ; trunk_s <- segema*trunk_s + (1-segema)*p, BN running stats included
        (copy_parameters keeps every non-0-dim state_dict entry). capgan: whole generator, parameters only
        (fedlab serialize_model, capgan.py:170-175)."""
        k = self.k
        if self.algo in ("mdgan", "acgan"):
            return  # these scripts have no Cloud
        if self.algo in ("cglgan", "mixed") and k.cloud_mode == "as_written":
            return  # keys '0.weight' never match 'model.0.weight': load_state_dict(strict=False) loads nothing
        with torch.no_grad():
            bank = self.G.trunk    # cglgan / mixed: net_g.model; capgan: the whole (single-path) generator
            assert bank.rows == self.S
            bufs = [(bank.params.detach(), bank.lay.ld)]
            if bank.lay.n_stats and self.algo in ("cglgan", "mixed"):
                bufs.append((bank.stats, bank.lay.ld_stats))   # dict form carries running_mean / running_var
            for buf, ld in bufs:
                g = torch.empty(ld, device=self.device)
                _wsum(self.A, None, buf, ld, g, self.comm)
                _bcast(None, k.segema, g, buf, ld, rows_n=bank.rows)

    def share_discriminators(self):
        """Neighbour-D share every E rounds (commented out in the shipped scripts, README.md:26):
        "group_mean": every client of a server ends with the mean D of that server's clients
        (ACGAN/MNIST/acgan.py:240-263 fixed point == CGLGAN/2DMG/main.py:171-179);
        "swap": the server shuffles its clients' Ds (MDGAN/MNIST/mdgan.py:158-164,258-262)."""
        row_ptr, col = [0], []
        for s, cl in enumerate(self.client_list):
            if self.k.d_share == "swap":
                order = list(range(len(cl)))
                self.swap_rd[s].shuffle(order)          # the shuffle of a list depends on its length only
                for j, c in enumerate(cl):              # clients are numbered server by server: row c == position in col
                    col.append(cl[order[j]])
                    row_ptr.append(len(col))
            else:
                for c in cl:
                    col += cl                           # `p += d` over the group in client order, then `p /= len`
                    row_ptr.append(len(col))
        row_ptr, col = torch.tensor(row_ptr, dtype=torch.int32), torch.tensor(col, dtype=torch.int32)
        if self.k.d_share == "swap":
            self.bank.mix((row_ptr, col, torch.ones(col.numel())))      # a permutation: exact copies
        else:
            self.bank.mix((row_ptr, col, None))                         # row means (sum, then one division)


class MDSingleServerSim(MDStyleSim):
    """ONE edge server whose clients are dealt over the ranks -- MD-GAN as shipped (MDGAN/MNIST/mdgan.py:35-36:
    num_servers = 1) on several GPUs. Every rank keeps a replica of the generator and the discriminators of its
    own block of clients; per round the ranks exchange
      * the clients' G losses (all-gather of N floats: the weighting of SURVEY 3.4 needs all of them), and
      * sum_i w_i dLoss_i/dXg, the gradient of the server objective with respect to the shared batch Xg ([B, d] floats,
        one all-reduce over NVLink) -- what F_max.backward() accumulates into Xg in the reference (mdgan.py:203-204),
    and every rank takes the identical generator step (same kernels, same inputs: the replicas stay bit-identical).
    z_d / z_g must be the same on every rank (default: a device generator seeded alike on all ranks).
    Single-path generators only (mdgan, acgan, capgan, capgan_copy, cglgan with iid == 0)."""

    def __init__(self, algo, knobs: Knobs, part_sizes=None, device="cuda", comm=None, rank=0, world=1, z_seed=20211212):
        from .dist import shard_range
        assert knobs.num_servers == 1, "MDSingleServerSim shards the clients of ONE server"
        self.N_total = knobs.num_workers
        self.rank, self.world = rank, world
        self.lo, self.hi = shard_range(self.N_total, world, rank)
        assert (self.hi - self.lo) * world == self.N_total, "the clients must divide evenly over the ranks (all-gather of the losses)"
        sizes = [1] * self.N_total if part_sizes is None else list(part_sizes)
        local = replace(knobs, num_workers=self.hi - self.lo)
        super().__init__(algo, local, part_sizes=sizes[self.lo:self.hi], device=device, comm=None)
        assert not self.multi_head, "one head per client would shard the generator itself: single-path generators only"
        self.xcomm = comm                                   # dist.ShardComm (None: one rank)
        full = torch.tensor(sizes, dtype=torch.float32).view(1, self.N_total)
        self.data_len = full.sum(1)
        self.beta = (full / self.data_len.unsqueeze(1)).to(self.device)      # shares over ALL clients of the server
        self.A = torch.ones(1, device=self.device)
        self._zgen = torch.Generator(device=self.device)
        self._zgen.manual_seed(z_seed)

    def round(self, real, n_real=None, z_d=None, z_g=None):
        """real [epoch, N_local, B, d] (this rank's clients). Returns the G losses of ALL clients [1, N_total]."""
        import torch.distributed as dist
        k, B, d = self.k, self.B, self.d
        if real.dim() == 3:
            real = real.unsqueeze(0)
        if n_real is not None and n_real.dim() == 1:
            n_real = n_real.unsqueeze(0)
        if self._share_due():
            self.share_discriminators()
        if z_d is None:
            z_d = torch.randn(1, B, 100, device=self.device, generator=self._zgen)
        if z_g is None:
            z_g = torch.randn(1, B, 100, device=self.device, generator=self._zgen)
        G = self.G
        Xd = G(z_d).reshape(1, B, d)
        Xg = G(z_g).reshape(1, B, d)
        for e in range(real.shape[0]):
            self.last_d_loss = self.bank.d_step(real[e], Xd, n_real=None if n_real is None else n_real[e],
                                                fake_idx=self.server_of)
        loss_local, dxg = self.bank.g_loss_raw(Xg, xg_idx=self.server_of)
        if self.world > 1:
            loss = torch.empty(self.N_total, device=self.device)
            dist.all_gather_into_tensor(loss, loss_local.contiguous())
        else:
            loss = loss_local
        loss = loss.view(1, self.N_total)
        w = self._server_weights(loss)                      # identical on every rank (replicated Lambda, all losses)
        w_local = w[0, self.lo:self.hi].contiguous().float()
        dy = torch.empty(1, B, d, device=self.device)
        abi.check(abi.lib.cgl_dxg_reduce(1, abi.ptr(self.srv_ptr), None, abi.ptr(w_local), abi.ptr(dxg), B * d, abi.ptr(dy),
                                         _stream()))
        if self.xcomm is not None:
            self.xcomm.allreduce_(dy)                       # the one data-path collective: [B, d] floats
        G.backward_step(dy.view(1, 1, B, d) if self.n_heads else dy)
        self.t += 1
        return loss

    def _cloud_due(self):
        return False    # a single server: the Cloud average of one generator is that generator

    def share_discriminators(self):
        """swap: the server's shuffle runs over ALL its clients (same seeded generator on every rank), every rank
        all-gathers the discriminator rows and keeps the ones dealt to its clients; group mean: local sum, all-reduce,
        one division (the order of the additions differs from the single-process loop by the partial sums)."""
        import torch.distributed as dist
        bank, N = self.bank, self.N_total
        if self.k.d_share == "swap":
            order = list(range(N))
            self.swap_rd[0].shuffle(order)
            if self.world > 1:
                full = torch.empty(N, bank.ld, device=self.device)
                dist.all_gather_into_tensor(full, bank.params.contiguous())
            else:
                full = bank.params.clone()
            take = torch.tensor(order[self.lo:self.hi], device=self.device, dtype=torch.long)
            bank.params.copy_(full.index_select(0, take))
        else:
            g = torch.empty(bank.ld, device=self.device)
            ones = torch.ones(bank.C, device=self.device)
            _wsum(ones, None, bank.params, bank.ld, g, self.xcomm)
            g.div_(float(N))
            _bcast(None, 0.0, g, bank.params, bank.ld, rows_n=bank.C)


def _wsum(w, rows, buf, ld, out, comm):
    w = w.contiguous()
    if comm is None:
        abi.check(abi.lib.cgl_wsum(w.numel(), ld, abi.ptr(w), abi.ptr(rows), abi.ptr(buf), ld, abi.ptr(out), _stream()))
    else:
        abi.check(abi.lib.cgl_mix_allreduce(comm.handle, w.numel(), ld, abi.ptr(w), abi.ptr(rows), abi.ptr(buf), ld,
                                            abi.ptr(out), _stream()))


def _bcast(rows, sigma, g, buf, ld, rows_n=None):
    R = rows.numel() if rows is not None else rows_n
    abi.check(abi.lib.cgl_bcast_mix(R, ld, abi.ptr(rows), float(sigma), abi.ptr(g), abi.ptr(buf), ld, _stream()))


class FLStyleSim:
    """FL-GAN: every client owns G and D; the server averages both every round.
    Worker.run/train FLGAN/MNIST/flgan.py:211-270, Server.run :134-163. Adam state stays with the client and
    is neither reset nor averaged when the parameters are overwritten (optimizers built once, :217-218)."""

    def __init__(self, knobs: Knobs, device="cuda", comm=None, weights=None):
        k = knobs
        self.k = k
        self.device = torch.device(device)
        self.C, self.B = k.num_workers, k.batch_size
        d = 1
        for s in k.img_shape:
            d *= s
        self.d = d
        self.G = StackedGenerator(k.img_shape, self.C, 0, device=self.device, lr=k.lr_g, b1=k.b1, b2=k.b2)
        self.bank = ClientBank(abi.ARCH_D_2D if d == 2 else abi.ARCH_D_MNIST1, self.C, self.B, device=self.device,
                               loss_kind=abi.LOSS_BCE, lr=k.lr_d, b1=k.b1, b2=k.b2)
        n_total = k.num_workers if comm is None else k.num_workers * comm.world
        self.n_total = n_total
        # p[key] += paras[key] / len(client_list)  (flgan.py:151-158): cgl_wsum_div; weights override = FeGAN's softmax(sk)
        self.w = None if weights is None else torch.as_tensor(weights, dtype=torch.float32).to(self.device)
        self.comm = comm
        self.t = 0
        self._ws = None

    def load_global(self, g_module, d_module):
        """Round-0 state: the server's initial net_g / net_d copied to every client (flgan.py:139-146)."""
        self.G.load_modules([g_module] * self.C)
        self.bank.load_modules([d_module] * self.C)

    def local_minibatch(self, real, n_real=None, z_d=None, z_g=None, client_ids=None):
        """One D step + one G step on every client (flgan.py:251-269), or on the clients listed in client_ids
        (int32 device tensor: FeGAN's group of the round; real / z then hold one entry per listed client).
        One call of the C ABI: cgl_fl_step."""
        C_, B = (self.C if client_ids is None else client_ids.numel()), self.B
        if C_ == 0:
            return torch.empty(0, device=self.device), torch.empty(0, device=self.device)
        if z_d is None:
            z_d = torch.randn(C_, B, 100, device=self.device)
        if z_g is None:
            z_g = torch.randn(C_, B, 100, device=self.device)
        G, bank = self.G, self.bank
        tb = G.trunk
        ids = None if client_ids is None else client_ids.to(device=self.device, dtype=torch.int32).contiguous()
        z_d = z_d.reshape(C_, B, 100).contiguous().float()
        z_g = z_g.reshape(C_, B, 100).contiguous().float()
        real = real.reshape(C_, B, self.d).contiguous()
        n_real = None if n_real is None else n_real.to(device=self.device, dtype=torch.int32).contiguous()
        nbytes = abi.lib.cgl_fl_step_workspace_bytes(ctypes.byref(tb.desc), ctypes.byref(bank.desc), C_, B)
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        d_loss, g_loss = torch.empty(C_, device=self.device), torch.empty(C_, device=self.device)
        abi.check(abi.lib.cgl_fl_step(
            ctypes.byref(tb.desc), ctypes.byref(bank.desc), C_, abi.ptr(tb.params), abi.ptr(tb.adam_m), abi.ptr(tb.adam_v),
            tb.lay.ld, abi.ptr(tb.step), abi.ptr(tb.stats), tb.lay.ld_stats, abi.ptr(bank.params), abi.ptr(bank.adam_m),
            abi.ptr(bank.adam_v), bank.ld, abi.ptr(bank.step), abi.ptr(ids), abi.ptr(z_d), abi.ptr(z_g), abi.ptr(real),
            abi.ptr(n_real), B, ctypes.byref(bank.cfg), ctypes.byref(G.cfg), abi.ptr(d_loss), abi.ptr(g_loss),
            abi.ptr(self._ws), self._ws.numel(), _stream()))
        return d_loss, g_loss

    def local_epochs(self, parts, epoch=1, z_fn=None, client_ids=None):
        """FL Worker.train as the MNIST scripts run it (FLGAN/MNIST/flgan.py:249-250, fegan.py:282-283): `epoch` FULL passes
        `for imgs in DataLoader(dataset, batch_size)` over the client's own partition, unshuffled, the last batch short.
        parts: data.ResidentPartitions (the dataset resident in HBM, one list of row ids per client). Clients with fewer
        batches simply stop earlier: minibatch j runs on the clients that still have a j-th batch.
        z_fn(n) -> (z_d, z_g) [n, B, 100] injects the noise (default: torch.randn on the device).
        Returns the number of client-minibatches done."""
        ids_all = list(range(self.C)) if client_ids is None else [int(c) for c in client_ids]
        B = self.B
        nb = {c: (len(parts.parts[c]) + B - 1) // B for c in ids_all}
        done = 0
        for _ in range(epoch):
            for j in range(max(nb.values()) if nb else 0):
                active = [c for c in ids_all if nb[c] > j]
                idx = torch.full((len(active), B), -1, dtype=torch.int64)
                n = torch.empty(len(active), dtype=torch.int32)
                for a, c in enumerate(active):
                    rows = parts.parts[c][j * B:(j + 1) * B]
                    idx[a, :rows.numel()] = rows
                    n[a] = rows.numel()
                real = parts.gather(idx)
                z_d, z_g = z_fn(len(active)) if z_fn else (None, None)
                cid = None if len(active) == self.C and client_ids is None else torch.tensor(active, dtype=torch.int32)
                self.local_minibatch(real, n.to(self.device), z_d, z_g, client_ids=cid)
                done += len(active)
        return done

    def aggregate(self):
        """Server.run: uniform (or weighted) average of every client's G and D, loaded back into every
        client (flgan.py:143-163,220-223). BN running stats take part (dict form, copy_parameters)."""
        with torch.no_grad():
            bufs = [(self.bank.params, self.bank.ld), (self.G.trunk.params.detach(), self.G.trunk.lay.ld)]
            if self.G.trunk.lay.n_stats:
                bufs.append((self.G.trunk.stats, self.G.trunk.lay.ld_stats))
            for buf, ld in bufs:
                g = torch.empty(ld, device=self.device)
                if self.w is not None:
                    _wsum(self.w, None, buf, ld, g, self.comm)
                else:       # every term divided by the client count, summed in client order (bit-exact on one GPU)
                    abi.check(abi.lib.cgl_wsum_div(self.C, ld, float(self.n_total), 0, None, abi.ptr(buf), ld, abi.ptr(g),
                                                   _stream()))
                    if self.comm is not None:
                        self.comm.allreduce_(g)
                _bcast(None, 0.0, g, buf, ld, rows_n=self.C)
        self.t += 1


class FeGANSim(FLStyleSim):
    """FeGAN (fegan.py:125-165, 220-303): every round ONE group of clients (partition.init_groups, frac_workers of
    the population) receives the global G and D -- parameters only: SerializationTool.deserialize_model leaves
    BatchNorm running statistics and the Adam state with the client -- trains locally like an FL-GAN client, and
    the server replaces the global vectors by fedavg_aggregate over the group with weights softmax(sk), sk the
    clients' KL scores (fegan.py:142-146,163-164).

    Several GPUs (comm given): the POPULATION is dealt over the ranks in contiguous blocks -- a client's Adam moments and
    BatchNorm statistics persist between the rounds it takes part in, so its state stays with its owner -- and a round's
    group is served by the owners of its members: no client state ever moves. The only exchange is the all-reduce of the two
    weighted partial sums (weights normalised over the WHOLE group), after which every rank holds the new global vectors.
    `knobs.num_workers` is the population size; ids / real / z of a round are those of the members this rank owns."""

    def __init__(self, knobs, sk, groups, device="cuda", comm=None, rank=0, world=1):
        from .dist import shard_range
        self.N_pop = knobs.num_workers
        self.rank, self.world = rank, world
        self.lo, self.hi = shard_range(self.N_pop, world, rank)
        super().__init__(replace(knobs, num_workers=self.hi - self.lo), device=device)
        self.xcomm = comm
        self.sk = torch.as_tensor(sk, dtype=torch.float32)
        self.groups = [list(g) for g in groups]
        self.p_g = torch.zeros(self.G.trunk.lay.ld, device=self.device)
        self.p_d = torch.zeros(self.bank.ld, device=self.device)

    def load_global(self, g_module, d_module):
        """p_g / p_d = serialize_model of the server's fresh networks (fegan.py:133-134)."""
        from .layout import flatten_params
        fg, fd = flatten_params(g_module).float(), flatten_params(d_module).float()
        self.p_g.zero_()
        self.p_g[:fg.numel()].copy_(fg)
        self.p_d.zero_()
        self.p_d[:fd.numel()].copy_(fd)

    def begin_round(self):
        """-> (members, ids): the members of the round's group this rank owns (global client numbers, in group order) and
        their local rows; the global vectors are loaded into those rows."""
        group = self.groups[self.t % len(self.groups)]
        mine = [c for c in group if self.lo <= c < self.hi]
        ids = torch.tensor([c - self.lo for c in mine], dtype=torch.int32, device=self.device)
        if mine:
            _bcast(ids, 0.0, self.p_d, self.bank.params, self.bank.ld)
            _bcast(ids, 0.0, self.p_g, self.G.trunk.params, self.G.trunk.lay.ld)
        return mine, ids

    def end_round(self, mine, ids):
        group = self.groups[self.t % len(self.groups)]
        w_all = torch.exp(self.sk[group])                  # weight = exp(sk); weight /= weight.sum()  (over the whole group)
        w_all = w_all / w_all.sum()
        pos = {c: j for j, c in enumerate(group)}
        w = w_all[[pos[c] for c in mine]].to(self.device) if mine else torch.zeros(0, device=self.device)
        for buf, ld, out in ((self.bank.params, self.bank.ld, self.p_d),
                             (self.G.trunk.params.detach(), self.G.trunk.lay.ld, self.p_g)):
            if mine:
                _wsum(w, ids, buf, ld, out, None)
            else:
                out.zero_()
            if self.xcomm is not None:
                self.xcomm.allreduce_(out)
        self.t += 1
