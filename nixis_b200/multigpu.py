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
from .partition import RankPlan, build_rank_plan, build_rank_plan_local, exchange_halo_torch, round_up, vertex_ranges, TILE
from .pipeline import Collective, assemble_heights, N_INIT_ROUGH, N_INIT_STRENGTH, N_ROUGHNESS, N_PERSISTENCE
from .util import DeviceMesh

RAIN_AMOUNT = 0.3 / 320


def _ptr_array(values):
    return (C.c_void_p * len(values))(*[C.c_void_p(int(v)) for v in values])


def _i64_array(values):
    return (C.c_int64 * len(values))(*[int(v) for v in values])


class _RawCuda:
    """A cudaMalloc'ed range as a torch-importable object (__cuda_array_interface__)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (int(n_floats),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


class PeerMemory:
    """One float32 buffer per rank that every rank of the group can store to from its kernels.
    kind "symm": torch.distributed._symmetric_memory (plumbing); kind "ipc": cudaMalloc + CUDA IPC
    handles exchanged through the process group (csrc/nxb_halo.cu nxb_peer_*), which also works when
    several ranks share one device.  `ptrs[r]` is rank r's buffer as seen from this process."""

    def __init__(self, n_floats, device, group, kind):
        self.kind, self.group = kind, group
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        self._opened = []
        self._own = None
        if kind == "symm":
            import torch.distributed._symmetric_memory as symm
            buf = symm.empty(int(n_floats), dtype=torch.float32, device=device)
            self._hdl = symm.rendezvous(buf, group if group is not None else dist.group.WORLD)
            buf.zero_()
            self.buf = buf
            self.ptrs = [int(p) for p in self._hdl.buffer_ptrs]
        elif kind == "ipc":
            ptr = C.c_void_p()
            handle = (C.c_ubyte * 64)()
            _lib.call("nxb_peer_alloc", int(n_floats) * 4, C.byref(ptr), handle)
            self._own = ptr.value
            self._raw = _RawCuda(ptr.value, n_floats)
            self.buf = torch.as_tensor(self._raw, device=device)
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle), group=group)
            self.ptrs = []
            for r in range(world):
                if r == rank:
                    self.ptrs.append(self._own)
                    continue
                p = C.c_void_p()
                _lib.call("nxb_peer_open", (C.c_ubyte * 64).from_buffer_copy(handles[r]), C.byref(p))
                self._opened.append(p.value)
                self.ptrs.append(p.value)
        else:
            raise ValueError(kind)

    def close(self):
        """Unmap the peers' buffers and free the own one (ipc kind; collective: everybody calls it)."""
        if self.kind != "ipc" or self._own is None:
            return
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for p in self._opened:
            _lib.call("nxb_peer_close", C.c_void_p(p))
        self._opened = []
        dist.barrier(group=self.group)
        self.buf = None
        _lib.call("nxb_peer_free", C.c_void_p(self._own))
        self._own = None


# Symmetric-memory buffers of closed ShardedErosion objects, kept for the next one of the same size: tearing a
# torch symmetric-memory allocation down costs 0.1 - 1 s (measured, tools/e2e_probe.py), which a caller of the
# numpy API would pay on EVERY erode_terrain3 call.  Every rank runs the same sequence of constructions with the
# same (all-reduced) size, so all ranks hit or miss together.
_PEER_POOL = {}


def _is_gloo(group):
    try:
        return dist.get_backend(group) == "gloo"
    except Exception:
        return False


MAX_PEERS = 8       # csrc/nxb_erosion.cu ERO_MAX_PEERS, csrc/nxb_halo.cu HALO_MAX_PEERS
MAX_FLAGS = 64      # flag slots behind the state buffers (one per source rank)
SEND_SCAN = 8       # csrc/nxb_erosion.cu ERO_SEND_SCAN


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
        cap_t = torch.tensor([plan.capacity], dtype=torch.int64, device="cpu" if _is_gloo(group) else dev)
        if self.world > 1:
            dist.all_reduce(cap_t, op=dist.ReduceOp.MAX, group=group)
        self.cap = int(cap_t.item())
        self.send_peers = sorted(plan.send_idx)
        self.recv_peers = sorted(plan.recv_slice)
        transport = transport if self.world > 1 else "none"
        if transport in ("nvlink", "fused"):
            # the kernels hold at most MAX_PEERS peer pointers and the flag array has MAX_FLAGS slots:
            # beyond that (world > 9 can touch more than 8 peers through the mesh skeleton) use p2p
            too_many = torch.tensor([int(len(self.send_peers) > MAX_PEERS or len(self.recv_peers) > MAX_PEERS
                                         or self.world > MAX_FLAGS)], dtype=torch.int64, device=cap_t.device)
            dist.all_reduce(too_many, op=dist.ReduceOp.MAX, group=group)
            if int(too_many.item()):
                transport = "p2p"
        self.sweeps = 0                     # total sweeps since creation (flag values)
        n_state = 4 * self.cap              # hA wA hB wB
        self.peer_mem = None
        if transport in ("nvlink", "fused"):
            kinds = [os.environ.get("NXB_PEER_MEM")] if os.environ.get("NXB_PEER_MEM") else \
                    (["ipc"] if _is_gloo(group) else ["symm", "ipc"])
            err = None
            self._pool_key = ("symm", n_state + MAX_FLAGS, str(dev), id(group))
            pooled = _PEER_POOL.pop(self._pool_key, None) if kinds[0] == "symm" else None
            if pooled is not None:
                pooled.buf.zero_()                  # flags and state of the previous user; everybody zeroes before the barrier below
                self.peer_mem = pooled
            for kind in ([] if pooled is not None else kinds):
                try:
                    self.peer_mem = PeerMemory(n_state + MAX_FLAGS, dev, group, kind)
                    break
                except Exception as exc:            # noqa: BLE001 -- fall through to the next mechanism
                    err = exc
            ok = torch.tensor([int(self.peer_mem is not None)], dtype=torch.int64, device=cap_t.device)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if not int(ok.item()):
                if self.peer_mem is not None:
                    self.peer_mem = None
                if os.environ.get("NXB_HALO_STRICT"):
                    raise RuntimeError(f"peer-mapped memory unavailable on rank {self.rank}: {err}")
                transport = "p2p"                   # NCCL / gloo point-to-point: the documented fallback
        self.transport = transport
        if self.peer_mem is not None:
            self._buf = self.peer_mem.buf
            self._peer_base = self.peer_mem.ptrs
        else:
            self._buf = torch.zeros(n_state + MAX_FLAGS, dtype=torch.float32, device=dev)
            self._peer_base = None
        c = self.cap
        # two buffer sets of interleaved {height, water} pairs, float32 [cap, 2] each
        self.hw = [self._buf[0:2 * c].view(c, 2), self._buf[2 * c:4 * c].view(c, 2)]
        self.sed = [torch.zeros(c, dtype=torch.float32, device=dev), torch.zeros(c, dtype=torch.float32, device=dev)]
        self.flags = self._buf[4 * c:4 * c + MAX_FLAGS].view(torch.int32)      # one uint32 per source rank
        self.ticket = torch.zeros(4, dtype=torch.int32, device=dev)
        self.cur = 0
        self._pending = False               # a publish has been issued whose incoming flags were not awaited yet
        # concatenated send list, peer after peer
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
            self._peer_arrays = {w: self._peer_ptrs(w) for w in (0, 1)}
        if self.world > 1:
            self._barrier()

    def _barrier(self):
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def close(self):
        if self.peer_mem is not None:
            if self.peer_mem.kind == "symm":
                if self._pending:
                    self._await()               # nobody may still be writing into this buffer
                self._barrier()
            self.hw = self.flags = self._buf = None
            if self.peer_mem.kind == "symm":
                _PEER_POOL.clear()              # keep one buffer
                _PEER_POOL[self._pool_key] = self.peer_mem
            else:
                self.peer_mem.close()
            self.peer_mem = None

    def _build_send_table(self):
        """Send entries {dst, vertex-in-tile, peer slot} grouped by tile (EroSendEntry); each tile's range
        of the list goes into its descriptor (send0, send1), from where the producer warp hands it to
        the consumers with the rest of the stage header."""
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
        # a tile is SPARSE when it has at most SEND_SCAN entries and no vertex occurs more than twice: its
        # threads then find their own entries by scanning the list (csrc/nxb_erosion.cu); other tiles are DENSE
        dense = counts > SEND_SCAN
        if vs:
            uv, mult = torch.unique(v, return_counts=True)
            over = torch.zeros(n_tiles, dtype=torch.bool, device=dev)
            over[(uv[mult > 2] // TILE)] = True
            dense |= over
            assert int(dst.max().item()) < (1 << 28), "peer slot index does not fit the packed send word"
        if n_tiles:
            desc = self.tile_plan.descriptors()
            e0 = ptr[:-1]
            desc[:, rt.ERO_DW_SEND] = torch.where(dense, -1 - e0, e0).to(torch.int32)
            desc[:, rt.ERO_DW_SEND + 1] = ptr[1:].to(torch.int32)
        self.n_send_tiles = int((counts > 0).sum().item())
        self.n_dense_send_tiles = int((dense & (counts > 0)).sum().item())

    # ------------------------------------------------------------------------------------
    def load(self, heights_own):
        """Start a run: own heights in, water / sediment zero (erosion.py:177-178), halos filled."""
        if self._pending:
            self._await()                   # drain the previous run's last incoming halo
        self.sweeps += 1                    # flag values of the new run never collide with the old run's
        hw = self.hw[self.cur]
        hw.zero_()
        hw[: self.plan.n_own, 0].copy_(heights_own[: self.plan.n_own])
        self.sed[0].zero_(); self.sed[1].zero_()
        self.hw[1 - self.cur].zero_()
        if self.world > 1:
            self._barrier()                 # nobody still reads / writes the old run's buffers
        self._publish(self.cur)

    def _publish(self, which):
        """Send this rank's boundary values of buffer set `which`; flag value = sweeps + 1."""
        if self.world == 1:
            return
        self._pending = True
        hw = self.hw[which]
        if self.transport in ("nvlink", "fused"):
            if not self.send_peers:
                return
            ph, pf = self._peer_ptrs(which)
            _lib.call("nxb_halo_put_f32", rt._ptr(hw), rt._ptr(self.send_idx), len(self.send_peers),
                      ph, pf, self._dst_off, self._src_begin, self._count,
                      C.c_uint32(self.sweeps + 1), rt._ptr(self.ticket), rt._stream())
        else:
            exchange_halo_torch(self.plan, [hw[:, 0], hw[:, 1]], group=self.group)

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
        """(peers' hw buffer of set `which`, peers' flag slot for this rank) as seen from this process"""
        c = self.cap
        off = (0 if which == 0 else 2 * c) * 4
        return (_ptr_array([self._peer_base[p] + off for p in self.send_peers]),
                _ptr_array([self._peer_base[p] + 4 * c * 4 + 4 * self.rank for p in self.send_peers]))

    def step(self, rain=RAIN_AMOUNT):
        self.run(1, rain)

    def run(self, n, rain=RAIN_AMOUNT):
        """n sweeps.  fused transport: ONE call, the loop (one sweep-with-exchange kernel per sweep that first
        awaits the peers' flags of the previous sweep, chained with programmatic dependent launch) is issued
        from C (nxb_erode3_run_comm_f32)."""
        if n <= 0:
            return
        if self.transport == "fused" and self.world > 1:
            tp = self.tile_plan
            d3 = tp.dist3_for(self.dist)
            in_c = self.wait_mode == "kernel"       # the flag waits are issued by the C loop too
            for n_call in ([n] if in_c else [1] * n):
                if not in_c:
                    self._await()
                a, b = (self.hw[self.cur], self.sed[self.cur]), (self.hw[1 - self.cur], self.sed[1 - self.cur])
                pa, pb = self._peer_arrays[self.cur], self._peer_arrays[1 - self.cur]
                n_wait = len(self.recv_peers) if in_c else 0
                _lib.call("nxb_erode3_run_comm_f32", rt._ptr(tp.mem), rt._ptr(tp.adj), rt._ptr(self.dist),
                          None if d3 is None else rt._ptr(d3),
                          rt._ptr(a[0]), rt._ptr(a[1]), rt._ptr(b[0]), rt._ptr(b[1]),
                          tp.n_own, C.c_float(rain), int(n_call),
                          rt._ptr(self.send_entries), len(self.send_peers), pa[0], pb[0], pa[1],
                          rt._ptr(self.flags), rt._ptr(self.recv_ranks), n_wait,
                          C.c_uint32(self.sweeps), rt._ptr(self.ticket), rt._stream(),
                          launches=int(n_call) * (2 if (n_wait and os.environ.get("NXB_ERO_WAIT_IN_SWEEP", "1") == "0") else 1))
                self._pending = True
                if n_call % 2:
                    self.cur = 1 - self.cur
                self.sweeps += n_call
            return
        for _ in range(n):
            src = (self.hw[self.cur], self.sed[self.cur])
            dst = (self.hw[1 - self.cur], self.sed[1 - self.cur])
            self._await()
            rt.erode3_step(self.tile_plan, self.dist, src, dst, rain)
            self.cur = 1 - self.cur
            self.sweeps += 1
            self._publish(self.cur)

    def finish(self):
        """Drain: wait for the last incoming halo so buffers may be reused."""
        self._await()

    @property
    def heights(self):
        return self.hw[self.cur][: self.plan.n_own, 0]

    @property
    def water(self):
        return self.hw[self.cur][: self.plan.n_own, 1]

    @property
    def sediment(self):
        return self.sed[self.cur][: self.plan.n_own]


class ShardedTerrain:
    """One rank's share of the whole hot path."""

    def __init__(self, k, seed=0, n_octaves=8, radius=1.0, transport="fused", group=None, noise_dim=3, w_scale=0.5):
        rt.require_cuda()
        assert noise_dim in (3, 4)
        self.noise_dim, self.w_scale, self.seed, self.n_octaves = int(noise_dim), float(w_scale), seed, int(n_octaves)
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
        import time
        t0 = time.perf_counter()
        # mesh shard: float64 positions of the own range only (closed form)
        self.xyz32, self.xyz64 = rt.mesh_points(self.k, self.begin, self.end, f32=(self.noise_dim == 4), f64=(self.noise_dim == 3),
                                                device=self.device)
        # neighbour rows of the OWN range only, straight from the closed-form triangle generator (no cell
        # array, no whole-mesh table on any rank); the halo plan comes from one exchange of id lists
        rows = rt.icosa_adj_rows(self.k, self.begin, self.end, device=self.device)
        self.plan = build_rank_plan_local(rows, self.rank, self.world, self.ranges, group=group)
        self.dist = rt.icosa_edge_lengths(self.k, rows, self.begin, self.end, self.radius)
        del rows
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        self.erosion = ShardedErosion(self.plan, self.dist, transport=transport, group=group)
        torch.cuda.synchronize()
        self.setup_ms = {"mesh_rows_plan_edges": (t1 - t0) * 1e3, "tile_plan_peer_memory": (time.perf_counter() - t1) * 1e3}

    def fbm(self, out=None, minmax=None):
        if self.noise_dim == 4:          # BASELINE configs[4]: 4-D noise, w = w_scale * frequency per octave
            return rt.fbm4(self.tables, self.xyz32, self.freq, self.amp, [self.w_scale * f for f in self.freq],
                           out=out, minmax=minmax)
        nr = [f / self.radius for f in self.freq]
        return rt.fbm3_pos64(self.tables, self.xyz64, self.radius, nr, self.amp, out=out, minmax=minmax)

    def heights(self, out=None):
        mm = rt.new_minmax(self.device)
        h = self.fbm(out=out, minmax=mm)
        return assemble_heights(h, coll=self.coll, mm=mm)


# ---------------------------------------------------------------------------------------------
def gather_own(t, ranges, group=None):
    """Own slices of every rank concatenated in rank order on every rank (all_gather, padded)."""
    sizes = [e - b for b, e in ranges]
    m = max(sizes)
    pad = torch.zeros(m, dtype=t.dtype, device=t.device)
    pad[: t.numel()] = t
    bufs = [torch.empty_like(pad) for _ in sizes]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])


def sharded_vs_single_gpu_check(terr, seed, n_octaves, sweeps=25):
    """the sharded pipeline (fBm + assembly + `sweeps` erosion sweeps with the halo
    exchange) against the SAME work on one GPU (rank 0 runs the whole planet next to its shard), h / w / s
    compared bit for bit.  Returns the verdict dict on every rank."""
    import hashlib
    from .pipeline import TerrainPipeline
    ero = terr.erosion
    h, _, lvl = terr.heights()
    ero.load(h)
    ero.run(sweeps - 2)
    ero.step(); ero.step()
    ero.finish()
    torch.cuda.synchronize()
    H = gather_own(ero.heights.contiguous(), terr.ranges, terr.group)
    W = gather_own(ero.water.contiguous(), terr.ranges, terr.group)
    S = gather_own(ero.sediment.contiguous(), terr.ranges, terr.group)
    verdict = torch.zeros(2, dtype=torch.int64, device=terr.device)
    sha = ""
    if terr.rank == 0:
        pipe = TerrainPipeline(terr.k, seed=seed, n_octaves=n_octaves, radius=terr.radius, noise_dim=terr.noise_dim, w_scale=terr.w_scale)
        pipe.build_mesh()
        h1, _, lvl1 = pipe.heights()
        st = pipe.erosion_state(h1)
        st.run(sweeps)
        torch.cuda.synchronize()
        same = torch.equal(H, st.heights) and torch.equal(W, st.water) and torch.equal(S, st.sediment) and lvl == lvl1
        verdict[0] = int(same)
        verdict[1] = int(torch.isfinite(H).all().item())
        sha = hashlib.sha1(H.cpu().numpy().tobytes()).hexdigest()
        del pipe, st, h1
    del H, W, S
    torch.cuda.empty_cache()
    dist.broadcast(verdict, 0, group=terr.group)
    return {"sweeps": sweeps, "bit_identical": bool(verdict[0].item()), "all_finite": bool(verdict[1].item()),
            "compared": "heights, water, sediment of all vertices + ocean level vs one GPU (torch.equal)",
            "sha1_heights": sha, "transport": ero.transport,
            "peer_memory": ero.peer_mem.kind if ero.peer_mem is not None else None}


def run_multi_gpu_bench(args, rank, world, local, emit=None):
    """bench.py --gpus N (N > 1): same step as the single-GPU arm, vertex range sharded over N ranks."""
    import json
    import statistics
    import sys
    import time
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench as B
    from . import shard

    k, n_oct, iters = args.division, args.octaves, args.iters
    transport = os.environ.get("NXB_HALO", "fused")
    # one-time process-level initialisation (NCCL communicators are created lazily by the first collective /
    # point-to-point call, symmetric memory by the first rendezvous: seconds) is kept out of the per-mesh setup time
    warm = ShardedTerrain(64, seed=args.seed, n_octaves=1, radius=1.0, transport=transport)
    warm.erosion.load(warm.heights()[0]); warm.erosion.run(2); warm.erosion.finish()
    torch.cuda.synchronize(); dist.barrier()
    warm.erosion.close()
    del warm
    _PEER_POOL.clear()
    torch.cuda.reset_peak_memory_stats()
    torch.cuda.synchronize(); dist.barrier()
    t_setup = time.perf_counter()
    terr = ShardedTerrain(k, seed=args.seed, n_octaves=n_oct, radius=1.0, transport=transport, noise_dim=args.noise_dim)
    torch.cuda.synchronize(); dist.barrier()
    setup_total_ms = (time.perf_counter() - t_setup) * 1e3
    peak_setup_gib = torch.cuda.max_memory_allocated() / 2 ** 30
    V, n_own = terr.V, terr.n_own
    ero = terr.erosion
    # ---- correctness first: sharded == one GPU, bit for bit, at THIS size, before anything is timed
    check = sharded_vs_single_gpu_check(terr, args.seed, n_oct, sweeps=25)
    if not check["bit_identical"]:
        if rank == 0:
            sys.stderr.write(f"mgpu_check FAILED: {json.dumps(check)}\n")
        raise SystemExit(3)
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
    # the same sweeps on finite data (first 100 sweeps of a run)
    h_fin, _, _ = terr.heights()
    fin = []
    for _ in range(3):
        ero.load(h_fin)
        torch.cuda.synchronize(); dist.barrier()
        a, b = ev(), ev()
        a.record(); ero.run(100); ero.finish(); b.record()
        torch.cuda.synchronize()
        fin.append(a.elapsed_time(b) / 100)
    fin_t = torch.tensor([min(fin)], dtype=torch.float64, device=terr.device)
    dist.all_reduce(fin_t, op=dist.ReduceOp.MAX)
    mem = torch.tensor([peak_setup_gib, torch.cuda.max_memory_allocated() / 2 ** 30], dtype=torch.float64, device=terr.device)
    dist.all_reduce(mem, op=dist.ReduceOp.MAX)
    tiles = torch.tensor([ero.tile_plan.n_tiles, ero.tile_plan.n_irregular, ero.tile_plan.n_affine, ero.tile_plan.n_affine3],
                         dtype=torch.int64, device=terr.device)
    dist.all_reduce(tiles)
    # ---- end to end through the reference-named numpy API: every rank passes ITS slices of the same
    # arrays to the same calls (bench.run_e2e, one definition for every N)
    e2e = None
    if not args.no_e2e and args.noise_dim == 3:       # the reference-named API has no 4-D fBm driver
        pts = torch.empty((n_own, 3), dtype=torch.float64, pin_memory=True)
        pts.copy_(terr.xyz64)
        nbr = torch.empty((n_own, 6), dtype=torch.int32, pin_memory=True)
        nbr.copy_(rt.icosa_adj_rows(k, terr.begin, terr.end))
        torch.cuda.synchronize()
        ero.close()                                   # the API calls build their own plan / peer memory
        shard.set_shard(terr.ranges)
        try:
            e2e = B.run_e2e(args, V, pts.numpy(), nbr.numpy(), terr.perm, terr.pgi, np, torch, sharded=True)
        finally:
            shard.clear_shard()
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
            "mgpu_check": check,
            "fbm_plus_assembly_ms": fbm_asm_ms,
            "erosion": {"value": V * iters / (ero_ms * 1e-3) / 1e6, "unit": "Mvert-iters/s", "ms": ero_ms,
                        "halo_transport": ero.transport, "max_halo_vertices_per_rank": int(halo[0].item()),
                        "max_sent_vertices_per_rank_per_sweep": int(halo[1].item()),
                        "nonfinite_heights_after_last_step": int(nonfinite.item()),
                        "finite_data": {"sweeps": 100, "ms_per_sweep": fin_t.item(), "value": V / (fin_t.item() * 1e-3) / 1e6,
                                        "unit": "Mvert-iters/s"}},
            "roofline": {"kernel": "erode3_plan_kernel (per GPU, incl. halo wait/put per sweep)", "bound": "hbm",
                         "achieved": ero_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": ero_gbs / hbm_peak,
                         "traffic": None, "peak_source": hbm_src,
                         "algorithmic_bytes_per_launch": B.BYTES_PER_VERT_ITER * n_own, "avg_launch_ms": ero_launch_ms,
                         "tiles_all_ranks": {"total": int(tiles[0]), "irregular": int(tiles[1]), "affine": int(tiles[2]),
                                             "affine_one_length_per_edge": int(tiles[3])}},
            "setup_ms": {"total_max_over_ranks_incl_barriers": setup_total_ms, **terr.setup_ms, "in_timed_step": False,
                         "note": "every rank builds ONLY its own vertex range (positions, neighbour rows from the closed-form "
                                 "triangle scan, halo plan by id-list exchange, edge lengths, tile plan, peer memory)"},
            "peak_device_memory_gib": {"after_setup_max_over_ranks": mem[0].item(), "whole_run_max_over_ranks": mem[1].item(),
                                       "note": "whole_run includes rank 0's single-GPU reference of mgpu_check"},
            "clocks": clocks, "gpu_launches": launches}
    if e2e is not None:
        line["e2e"] = e2e
    (emit or (lambda ln: print(json.dumps(ln))))(line)
