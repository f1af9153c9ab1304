"""One-process-per-GPU driver of the hot path (torch.distributed for the plumbing).

fBm / height assembly: every rank works on its own contiguous vertex range; the only exchange is a
handful of scalars (global min/max, power_rescale statistics) through all_reduce / all_gather.
Erosion: one halo exchange of boundary h / w per sweep (partition.py).  Two transports:
  * "nvlink": state buffers live in torch symmetric memory, the peers' addresses are mapped into
    this process, and hand-written kernels put the boundary values straight into the peers' halo
    slots and raise a flag (csrc/nxb_halo.cu) -- three kernel launches per sweep (wait, sweep, put),
    no NCCL call, no host synchronisation;
  * "p2p": torch.distributed batch_isend_irecv (NCCL on GPUs, gloo in the CPU tests) -- the baseline
    and fallback.
"""
import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from . import runtime as rt
from .partition import RankPlan, build_rank_plan, exchange_halo_torch, round_up, vertex_ranges, TILE
from .pipeline import Collective, assemble_heights, N_INIT_ROUGH, N_INIT_STRENGTH, N_ROUGHNESS, N_PERSISTENCE
from .util import DeviceMesh

RAIN_AMOUNT = 0.3 / 320


def _ptr_array(values):
    return (C.c_void_p * len(values))(*[C.c_void_p(int(v)) for v in values])


def _i64_array(values):
    return (C.c_int64 * len(values))(*[int(v) for v in values])


class ShardedErosion:
    """Erosion state of one rank: padded own range + halo slots, ping-pong (h, w, s)."""

    def __init__(self, plan: RankPlan, dist_f32, transport="nvlink", group=None):
        self.plan, self.group = plan, group
        self.rank, self.world = plan.rank, plan.world
        dev = plan.local_adj.device
        self.device = dev
        self.dist = dist_f32
        self.tile_plan = rt.ErosionPlan(plan.local_adj, capacity=plan.capacity)
        # symmetric sizes: every rank allocates the max capacity so that offsets agree
        cap_t = torch.tensor([plan.capacity], dtype=torch.int64, device=dev)
        if self.world > 1:
            dist.all_reduce(cap_t, op=dist.ReduceOp.MAX, group=group)
        self.cap = int(cap_t.item())
        self.transport = transport if self.world > 1 else "none"
        self.sweeps = 0                     # total sweeps since creation (flag values)
        self.flag_base = 0
        n_state = 4 * self.cap              # hA wA hB wB
        if self.transport == "nvlink":
            import torch.distributed._symmetric_memory as symm
            self._symm = symm
            try:
                buf = symm.empty(int(n_state + 64), dtype=torch.float32, device=dev)
            except Exception as exc:
                raise RuntimeError(f"symmetric memory allocation failed on rank {self.rank}: n={n_state + 64} "
                                   f"cap={self.cap} plan.capacity={plan.capacity} dev={dev}: {exc}") from exc
            self._hdl = symm.rendezvous(buf, group if group is not None else dist.group.WORLD)
            buf.zero_()
            self._buf = buf
            self._peer_base = [int(p) for p in self._hdl.buffer_ptrs]
        else:
            self._buf = torch.zeros(n_state + 64, dtype=torch.float32, device=dev)
            self._peer_base = None
        c = self.cap
        self.hw = [(self._buf[0:c], self._buf[c:2 * c]), (self._buf[2 * c:3 * c], self._buf[3 * c:4 * c])]
        self.sed = [torch.zeros(c, dtype=torch.float32, device=dev), torch.zeros(c, dtype=torch.float32, device=dev)]
        self.flags = self._buf[4 * c:4 * c + 64].view(torch.int32)      # one uint32 per source rank
        self.ticket = torch.zeros(4, dtype=torch.int32, device=dev)
        self.cur = 0
        self._pending = False               # a publish has been issued whose incoming flags were not awaited yet
        # concatenated send list, peer after peer
        self.send_peers = sorted(plan.send_idx)
        self.recv_peers = sorted(plan.recv_slice)
        if self.send_peers:
            self.send_idx = torch.cat([plan.send_idx[p] for p in self.send_peers]).contiguous()
        else:
            self.send_idx = torch.zeros(1, dtype=torch.int32, device=dev)
        counts = [int(plan.send_idx[p].numel()) for p in self.send_peers]
        self._count = _i64_array(counts)
        self._src_begin = _i64_array(np.concatenate([[0], np.cumsum(counts)[:-1]]) if counts else [])
        self._dst_off = _i64_array([plan.peer_n_own_pad[p] + plan.send_dst_offset[p] for p in self.send_peers])
        self.recv_ranks = torch.tensor(self.recv_peers if self.recv_peers else [0], dtype=torch.int32, device=dev)
        if self.world > 1:
            dist.barrier(group=group)

    # ------------------------------------------------------------------------------------
    def load(self, heights_own):
        """Start a run: own heights in, water / sediment zero (erosion.py:177-178), halos filled."""
        if self._pending:
            self._await()                   # drain the previous run's last incoming halo
        self.sweeps += 1                    # flag values of the new run never collide with the old run's
        h, w = self.hw[self.cur]
        h.zero_(); w.zero_()
        h[: self.plan.n_own].copy_(heights_own[: self.plan.n_own])
        self.sed[0].zero_(); self.sed[1].zero_()
        self.hw[1 - self.cur][0].zero_(); self.hw[1 - self.cur][1].zero_()
        if self.world > 1:
            torch.cuda.synchronize() if self.device.type == "cuda" else None
            dist.barrier(group=self.group)          # nobody still reads / writes the old run's buffers
        self._publish(self.cur)

    def _publish(self, which):
        """Send this rank's boundary values of buffer set `which`; flag value = sweeps + 1."""
        if self.world == 1:
            return
        self._pending = True
        h, w = self.hw[which]
        if self.transport == "nvlink":
            if not self.send_peers:
                return
            c = self.cap
            off_h = (0 if which == 0 else 2 * c) * 4
            off_w = (c if which == 0 else 3 * c) * 4
            ph = _ptr_array([self._peer_base[p] + off_h for p in self.send_peers])
            pw = _ptr_array([self._peer_base[p] + off_w for p in self.send_peers])
            pf = _ptr_array([self._peer_base[p] + 4 * c * 4 + 4 * self.rank for p in self.send_peers])
            _lib.call("nxb_halo_put_f32", rt._ptr(h), rt._ptr(w), rt._ptr(self.send_idx), len(self.send_peers),
                      ph, pw, pf, self._dst_off, self._src_begin, self._count,
                      C.c_uint32(self.sweeps + 1), rt._ptr(self.ticket), rt._stream())
        else:
            exchange_halo_torch(self.plan, [h, w], group=self.group)

    def _await(self):
        self._pending = False
        if self.world > 1 and self.transport == "nvlink" and self.recv_peers:
            _lib.call("nxb_halo_wait", rt._ptr(self.flags), rt._ptr(self.recv_ranks), len(self.recv_peers),
                      C.c_uint32(self.sweeps + 1), rt._stream())

    def step(self, rain=RAIN_AMOUNT):
        self._await()
        src = self.hw[self.cur] + (self.sed[self.cur],)
        dst = self.hw[1 - self.cur] + (self.sed[1 - self.cur],)
        rt.erode3_step(self.tile_plan, self.dist, src, dst, rain)
        self.cur = 1 - self.cur
        self.sweeps += 1
        self._publish(self.cur)

    def run(self, n, rain=RAIN_AMOUNT):
        for _ in range(n):
            self.step(rain)

    def finish(self):
        """Drain: wait for the last incoming halo so buffers may be reused."""
        self._await()

    @property
    def heights(self):
        return self.hw[self.cur][0][: self.plan.n_own]

    @property
    def water(self):
        return self.hw[self.cur][1][: self.plan.n_own]

    @property
    def sediment(self):
        return self.sed[self.cur][: self.plan.n_own]


class ShardedTerrain:
    """One rank's share of the whole hot path."""

    def __init__(self, k, seed=0, n_octaves=8, radius=1.0, transport="nvlink", group=None):
        rt.require_cuda()
        self.k, self.radius, self.group = int(k), float(radius), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.V = 10 * self.k ** 2 + 2
        self.perm, self.pgi = rt.init_perm(seed)
        self.tables = rt.tables_for(self.perm, self.pgi)
        self.freq, self.amp = rt.octave_schedule(n_octaves, N_INIT_ROUGH, N_INIT_STRENGTH, N_ROUGHNESS, N_PERSISTENCE)
        self.coll = Collective(group, distributed=True)
        self.ranges = vertex_ranges(self.V, self.world)
        self.begin, self.end = self.ranges[self.rank]
        self.n_own = self.end - self.begin
        # mesh shard (positions of the own range only)
        self.xyz, _ = rt.mesh_points(self.k, self.begin, self.end, device=self.device)
        # neighbour table: every rank builds the whole sorted table, then plans locally (partition.py)
        cells = rt.mesh_cells(self.k, device=self.device)
        unsorted = rt.adj_build(cells, self.V)
        del cells
        adj_global = rt.adj_sort(unsorted)
        del unsorted
        self.plan = build_rank_plan(adj_global, self.rank, self.world, self.ranges)
        self.dist = rt.icosa_edge_lengths(self.k, adj_global[self.begin:self.end].contiguous(), self.begin, self.end, self.radius)
        del adj_global
        torch.cuda.empty_cache()
        self.erosion = ShardedErosion(self.plan, self.dist, transport=transport, group=group)

    def fbm(self, out=None, minmax=None):
        return rt.fbm3(self.tables, self.xyz, self.freq, self.amp, out=out, minmax=minmax)

    def heights(self, out=None):
        mm = rt.new_minmax(self.device)
        h = self.fbm(out=out, minmax=mm)
        return assemble_heights(h, coll=self.coll, mm=mm)


# ---------------------------------------------------------------------------------------------
def run_multi_gpu_bench(args, rank, world, local):
    """bench.py --gpus N (N > 1): same step as the single-GPU arm, vertex range sharded over N ranks."""
    import json
    import statistics
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench as B

    k, n_oct, iters = args.division, args.octaves, args.iters
    transport = os.environ.get("NXB_HALO", "nvlink")
    terr = ShardedTerrain(k, seed=args.seed, n_octaves=n_oct, radius=1.0, transport=transport)
    V, n_own = terr.V, terr.n_own
    ero = terr.erosion
    h_buf = torch.empty(n_own, dtype=torch.float32, device=terr.device)
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(timed=None):
        e = [ev() for _ in range(4)]
        e[0].record()
        h, _, _ = terr.heights(out=h_buf)
        e[1].record()
        ero.load(h)
        e[2].record()
        ero.run(iters)
        ero.finish()
        e[3].record()
        if timed is not None:
            timed.append(e)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    sampler = B.ClockSampler(local)
    sampler.start()
    launches0 = _lib.launch_count
    timed = []
    t0, t1 = ev(), ev()
    torch.cuda.synchronize()
    dist.barrier()
    t0.record()
    for _ in range(args.steps):
        step(timed)
    t1.record()
    torch.cuda.synchronize()
    dist.barrier()
    clocks = sampler.stop()
    launches = _lib.launch_count - launches0
    ms = torch.tensor([t0.elapsed_time(t1) / args.steps,
                       statistics.mean(e[0].elapsed_time(e[1]) for e in timed),
                       statistics.mean(e[2].elapsed_time(e[3]) for e in timed)], dtype=torch.float64, device=terr.device)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_per_step, fbm_asm_ms, ero_ms = ms.tolist()
    halo = torch.tensor([terr.plan.n_halo, sum(int(v.numel()) for v in terr.plan.send_idx.values())],
                        dtype=torch.int64, device=terr.device)
    dist.all_reduce(halo, op=dist.ReduceOp.MAX)
    nonfinite = torch.tensor([int((~torch.isfinite(ero.heights)).sum().item())], dtype=torch.int64, device=terr.device)
    dist.all_reduce(nonfinite)
    if rank != 0:
        return
    value = V * (n_oct + iters) / (ms_per_step * 1e-3) / 1e6
    hbm_peak, hbm_src = B.measured_peaks()
    ero_launch_ms = ero_ms / iters
    # per-GPU roofline of the dominant kernel: this rank's share of the vertices per launch
    ero_gbs = B.BYTES_PER_VERT_ITER * n_own / (ero_launch_ms * 1e-3) / 1e9
    line = {"metric": B.METRIC, "value": value, "unit": B.UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": B.workload_config(args),
            "fbm_plus_assembly_ms": fbm_asm_ms,
            "erosion": {"value": V * iters / (ero_ms * 1e-3) / 1e6, "unit": "Mvert-iters/s", "ms": ero_ms,
                        "halo_transport": transport, "max_halo_vertices_per_rank": int(halo[0].item()),
                        "max_sent_vertices_per_rank_per_sweep": int(halo[1].item()),
                        "nonfinite_heights_after_last_step": int(nonfinite.item())},
            "roofline": {"kernel": "erode3_plan_kernel (per GPU, incl. halo wait/put per sweep)", "bound": "hbm",
                         "achieved": ero_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ero_gbs / hbm_peak,
                         "traffic": None, "peak_source": hbm_src,
                         "algorithmic_bytes_per_launch": B.BYTES_PER_VERT_ITER * n_own, "avg_launch_ms": ero_launch_ms},
            "clocks": clocks, "gpu_launches": launches}
    print(json.dumps(line))
