"""One-process-per-GPU driver of the hot path (torch.distributed for the plumbing).

fBm / height assembly: every rank works on its own contiguous vertex range; the only exchange is a
handful of scalars (global min/max, power_rescale statistics) through all_reduce / all_gather.
Erosion: one halo exchange of boundary h / w per sweep (partition.py).  Three transports:
  * "fused" (default): state buffers live in torch symmetric memory, the peers' addresses are mapped
    into this process, and the sweep kernel itself stores boundary results into the peers' halo slots
    over NVLink as it computes them and raises this rank's flag from its last CTA
    (csrc/nxb_erosion.cu, EroComm); a one-warp kernel waits for the peers' flags before the next
    sweep -- two launches per sweep, no NCCL call, no host synchronisation;
  * "nvlink": same memory, but separate wait / sweep / put kernels (csrc/nxb_halo.cu);
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

    def __init__(self, plan: RankPlan, dist_f32, transport="fused", group=None):
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
        if self.transport in ("nvlink", "fused"):
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
        self.ticket = torch.zeros(4 + 256, dtype=torch.int32, device=dev)
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
        self._recv_ranks_host = (C.c_int32 * max(1, len(self.recv_peers)))(*self.recv_peers)
        # "kernel": one-warp spin kernel; "stream": cuStreamWaitValue32 memory operations (no launch)
        self.wait_mode = os.environ.get("NXB_HALO_WAIT", "kernel")
        if self.transport == "fused":
            self._build_send_table()
        if self.world > 1:
            dist.barrier(group=group)

    def _build_send_table(self):
        """Per-tile CSR of {dst, vertex-in-tile, peer slot} for the fused sweep (EroSendEntry)."""
        plan, dev = self.plan, self.device
        n_tiles = (plan.n_own + TILE - 1) // TILE
        vs, dsts, slots = [], [], []
        for slot, p in enumerate(self.send_peers):
            idx = plan.send_idx[p].to(torch.int64)
            base = plan.peer_n_own_pad[p] + plan.send_dst_offset[p]
            vs.append(idx)
            dsts.append(base + torch.arange(idx.numel(), dtype=torch.int64, device=dev))
            slots.append(torch.full_like(idx, slot))
        if vs:
            v, dst, slot = torch.cat(vs), torch.cat(dsts), torch.cat(slots)
            order = torch.argsort(v // TILE, stable=True)
            v, dst, slot = v[order], dst[order], slot[order]
            counts = torch.bincount(v // TILE, minlength=n_tiles)
            packed = (dst & 0xFFFFFFFF) | ((v % TILE) << 32) | (slot << 48)       # little-endian {i32, u16, u16}
            self.send_entries = packed.contiguous()
        else:
            counts = torch.zeros(n_tiles, dtype=torch.int64, device=dev)
            self.send_entries = torch.zeros(1, dtype=torch.int64, device=dev)
        ptr = torch.zeros(n_tiles + 1, dtype=torch.int64, device=dev)
        ptr[1:] = torch.cumsum(counts, 0)
        self.send_ptr = ptr.to(torch.int32).contiguous()
        self._wait_rank = (C.c_int32 * max(1, len(self.recv_peers)))(*self.recv_peers)
        # processing order: tiles that read halo slots (a halo segment, or irregular) go last
        desc = self.tile_plan.mem[: n_tiles * 128].view(torch.int32).view(n_tiles, 32)
        NSEG = 8                      # nxb_erosion_plan.cuh: seg_start | seg_len | seg_off | nseg | irregular | ...
        seg_start, nseg, irregular = desc[:, 0:NSEG].to(torch.int64), desc[:, 2 * NSEG:2 * NSEG + 1], desc[:, 2 * NSEG + 1]
        live = torch.arange(NSEG, device=dev).unsqueeze(0) < nseg
        needs_halo = ((seg_start >= plan.n_own_pad) & live).any(dim=1) | (irregular != 0)
        self.tile_order = torch.argsort(needs_halo.to(torch.int8), stable=True).to(torch.int32).contiguous()
        self.n_halo_tiles = int(needs_halo.sum().item())
        # boundary set = tiles that send or read halo slots, FIRST (early flag mode)
        boundary = needs_halo | (counts > 0)
        self.boundary_order = torch.argsort((~boundary).to(torch.int8), stable=True).to(torch.int32).contiguous()
        self.n_boundary_tiles = int(boundary.sum().item())
        self.fused_mode = os.environ.get("NXB_FUSED_MODE", "inkernel" if os.environ.get("NXB_FUSED_INKERNEL_WAIT") else "sepwait")
        if self.fused_mode == "early" and (self.n_boundary_tiles == 0 or not self.send_peers):
            self.fused_mode = "sepwait"

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
        if self.transport in ("nvlink", "fused"):
            if not self.send_peers:
                return
            ph, pw, pf = self._peer_ptrs(which)
            _lib.call("nxb_halo_put_f32", rt._ptr(h), rt._ptr(w), rt._ptr(self.send_idx), len(self.send_peers),
                      ph, pw, pf, self._dst_off, self._src_begin, self._count,
                      C.c_uint32(self.sweeps + 1), rt._ptr(self.ticket), rt._stream())
        else:
            exchange_halo_torch(self.plan, [h, w], group=self.group)

    def _await(self):
        self._pending = False
        if self.world > 1 and self.transport in ("nvlink", "fused") and self.recv_peers:
            if self.wait_mode == "stream":
                _lib.call("nxb_halo_wait_stream", rt._ptr(self.flags), self._recv_ranks_host, len(self.recv_peers),
                          C.c_uint32(self.sweeps + 1), rt._stream())
            else:
                _lib.call("nxb_halo_wait", rt._ptr(self.flags), rt._ptr(self.recv_ranks), len(self.recv_peers),
                          C.c_uint32(self.sweeps + 1), rt._stream())

    def _peer_ptrs(self, which):
        c = self.cap
        off_h = (0 if which == 0 else 2 * c) * 4
        off_w = (c if which == 0 else 3 * c) * 4
        return (_ptr_array([self._peer_base[p] + off_h for p in self.send_peers]),
                _ptr_array([self._peer_base[p] + off_w for p in self.send_peers]),
                _ptr_array([self._peer_base[p] + 4 * c * 4 + 4 * self.rank for p in self.send_peers]))

    def step(self, rain=RAIN_AMOUNT):
        src = self.hw[self.cur] + (self.sed[self.cur],)
        dst = self.hw[1 - self.cur] + (self.sed[1 - self.cur],)
        if self.transport == "fused" and self.world > 1:
            # one kernel: wait (only where halo data is read) + sweep + put + flag
            ph, pw, pf = self._peer_ptrs(1 - self.cur)
            tp = self.tile_plan
            # The flag wait stays a tiny stream-ordered kernel: waiting INSIDE the sweep (n_wait > 0,
            # NXB_FUSED_INKERNEL_WAIT=1) works and is bit-identical, but measured bistable on 2 GPUs --
            # the ranks either stay in lockstep (333 us/sweep) or fall into a wait/compute alternation
            # (635 us/sweep); see profiles/r01_fused_wait_timeline.txt.
            n_wait, wait_target, order, n_early = 0, 0, None, 0
            if self.fused_mode == "early":
                # boundary tiles first, flags raised as soon as they are done, wait inside the kernel
                # (the peers' flags of the previous sweep went up a whole interior phase ago)
                n_wait, wait_target = len(self.recv_peers), self.sweeps + 1
                order, n_early = rt._ptr(self.boundary_order), self.n_boundary_tiles
            elif self.fused_mode == "inkernel":
                n_wait, wait_target = len(self.recv_peers), self.sweeps + 1
                order = rt._ptr(self.tile_order) if os.environ.get("NXB_FUSED_TILE_ORDER") else None
            else:
                self._await()
            d3 = tp.dist3_for(self.dist)
            _lib.call("nxb_erode3_plan_step_comm_f32", rt._ptr(tp.mem), rt._ptr(tp.adj), rt._ptr(self.dist),
                      None if d3 is None else rt._ptr(d3), rt._ptr(src[0]), rt._ptr(src[1]), rt._ptr(src[2]), rt._ptr(dst[0]), rt._ptr(dst[1]), rt._ptr(dst[2]),
                      tp.n_own, C.c_float(rain),
                      rt._ptr(self.send_ptr), rt._ptr(self.send_entries), len(self.send_peers), ph, pw, pf,
                      rt._ptr(self.flags), self._wait_rank, n_wait,
                      C.c_uint32(wait_target), C.c_uint32(self.sweeps + 2), self.plan.n_own_pad,
                      rt._ptr(self.ticket), order, n_early, rt._stream())
            self._pending = True
            self.cur = 1 - self.cur
            self.sweeps += 1
            return
        self._await()
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

    def __init__(self, k, seed=0, n_octaves=8, radius=1.0, transport="fused", group=None):
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
        # the skeleton's tiles are the irregular ones of the sweep planner (~6x a regular tile)
        self.ranges = vertex_ranges(self.V, self.world, front=12 + 30 * (self.k - 1),
                                    front_cost=float(os.environ.get("NXB_SKELETON_COST", "6")))
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

    def run_from_host(self, points_host, n_sweeps, out_host, world_radius=1.0):
        """End-to-end step of this rank with HOST buffers: `points_host` float64 [n_own,3] (this rank's
        slice of the reference's `points`, radius-scaled), `out_host` float64 [n_own] receives the
        eroded heights.  H2D of the positions and D2H of the result happen here, every call."""
        xyz = rt.xyz_from_f64(rt.upload(points_host), 1.0 / float(world_radius))
        mm = rt.new_minmax(self.device)
        h = rt.fbm3(self.tables, xyz, self.freq, self.amp, minmax=mm)
        h, _, _ = assemble_heights(h, coll=self.coll, mm=mm)
        self.erosion.load(h)
        self.erosion.run(n_sweeps)
        self.erosion.finish()
        return rt.download_f64(self.erosion.heights.contiguous(), out=out_host)


# ---------------------------------------------------------------------------------------------
def run_multi_gpu_bench(args, rank, world, local, emit=None):
    """bench.py --gpus N (N > 1): same step as the single-GPU arm, vertex range sharded over N ranks."""
    import json
    import statistics
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench as B

    k, n_oct, iters = args.division, args.octaves, args.iters
    transport = os.environ.get("NXB_HALO", "fused")
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

    sampler = B.ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    launches0 = _lib.launch_count
    timed = []
    t0, t1 = ev(), ev()
    torch.cuda.synchronize()
    dist.barrier()
    sampler.mark_begin()
    t0.record()
    for _ in range(args.steps):
        step(timed)
    t1.record()
    torch.cuda.synchronize()
    sampler.mark_end()
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
    # ---- end to end with host buffers: every rank uploads its slice of `points`, downloads its heights
    e2e = None
    if not args.no_e2e:
        import time
        pts = torch.empty((n_own, 3), dtype=torch.float64, pin_memory=True)
        pts.copy_(rt.mesh_points(k, terr.begin, terr.end, f32=False, f64=True)[1])
        out_h = torch.empty(n_own, dtype=torch.float64, pin_memory=True)
        pts_np, out_np = pts.numpy(), out_h.numpy()
        terr.run_from_host(pts_np, iters, out_np)            # warm-up
        torch.cuda.synchronize(); dist.barrier()
        reps = max(1, min(2, args.steps))
        t_0 = time.perf_counter()
        for _ in range(reps):
            terr.run_from_host(pts_np, iters, out_np)
        torch.cuda.synchronize(); dist.barrier()
        dt = torch.tensor([(time.perf_counter() - t_0) / reps], dtype=torch.float64, device=terr.device)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": V * (n_oct + iters) / dt.item() / 1e6, "unit": B.UNIT, "ms_per_step": dt.item() * 1e3,
               "h2d_bytes_per_step": int(V * 24), "d2h_bytes_per_step": int(V * 8),
               "api": "nixis_b200.multigpu.ShardedTerrain.run_from_host: every rank uploads its float64 slice of "
                      "`points` from pinned memory and downloads its float64 heights (bytes summed over ranks)"}
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
    if e2e is not None:
        line["e2e"] = e2e
    (emit or (lambda ln: print(json.dumps(ln))))(line)
