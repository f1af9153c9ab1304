"""Drop-in for the live half of the reference's `erosion` module (erosion.py) on B200.

Kept: `erode_terrain3` / `erosion_iteration3` (the variant nixis.py:410 calls) and
`erode_terrain1` / `erosion_iteration1` (the numerically stable variant).  Variants 2, 4, 5, 6
are dead, buggy or non-deterministic experiments in the reference and are
not provided.

State on the device is FP32 and ping-pongs between two buffer sets; there is no copy-back pass
(erosion.py:199-201, 274-277) and `water += rain` (erosion.py:182-183) is fused into the sweep.
numpy arguments are updated in place exactly where the reference updates them.
"""
import numpy as np
import torch

from . import runtime as rt
from . import shard
from .util import DeviceMesh

RAIN_AMOUNT = 0.3 / 320         # erosion.py:182


def _neighbors(neighbors):
    if isinstance(neighbors, torch.Tensor):
        return neighbors
    return rt.upload(np.ascontiguousarray(neighbors, dtype=np.int32))


def _edge_lengths(nodes, adj):
    """FP32 edge length of every adjacency slot, computed once in FP64 on the device
    (erosion.py:34-40, 227-229: distances between the undisplaced sphere positions)."""
    if isinstance(nodes, DeviceMesh):
        return rt.icosa_edge_lengths(nodes.k, adj, nodes.v_begin, nodes.v_begin + adj.shape[0], nodes.radius)
    if isinstance(nodes, torch.Tensor):
        n64 = nodes[:, :3].to(torch.float64).contiguous()
        return rt.edge_lengths(n64, adj)
    v = np.ascontiguousarray(nodes, dtype=np.float64)
    return rt.edge_lengths(rt.upload(v), adj)


class Erosion3State:
    """Device-resident state of erode_terrain3: tile plan, edge lengths, ({height, water}, sediment) x ping-pong.
    Heights and water are interleaved (float32 [capacity, 2]): a neighbour's two values are always read
    together, so they travel in one bulk copy / one 8-byte load.  Buffers are padded to whole 256-vertex tiles."""

    def __init__(self, nodes, adj, heights32, plan=None, dist=None):
        self.adj = adj
        self.n = adj.shape[0]
        self.plan = plan if plan is not None else rt.ErosionPlan(adj)
        self.dist = dist if dist is not None else _edge_lengths(nodes, adj)
        cap, dev = self.plan.capacity, adj.device
        hw = torch.zeros((cap, 2), dtype=rt.F32, device=dev)
        hw[: self.n, 0].copy_(heights32[: self.n])
        self.cur = (hw, torch.zeros(cap, dtype=rt.F32, device=dev))
        self.nxt = (torch.zeros((cap, 2), dtype=rt.F32, device=dev), torch.zeros(cap, dtype=rt.F32, device=dev))
        self.iterations = 0

    def reset(self, heights32):
        self.cur[0].zero_()
        self.cur[0][: self.n, 0].copy_(heights32[: self.n])
        self.cur[1].zero_()
        self.iterations = 0

    def step(self, rain=RAIN_AMOUNT):
        rt.erode3_step(self.plan, self.dist, self.cur, self.nxt, rain)
        self.cur, self.nxt = self.nxt, self.cur
        self.iterations += 1

    def run(self, num_iter, rain=RAIN_AMOUNT):
        """num_iter sweeps issued by the C-side loop (nxb_erode3_run_f32)."""
        if num_iter <= 0:
            return
        res = rt.erode3_run(self.plan, self.dist, self.cur, self.nxt, rain, num_iter)
        if res is self.nxt:
            self.cur, self.nxt = self.nxt, self.cur
        self.iterations += num_iter

    @property
    def heights(self):
        """float32 [n] view (stride 2) of the interleaved state."""
        return self.cur[0][: self.n, 0]

    @property
    def water(self):
        """Water AFTER the last sweep (the reference's `water` array at that point)."""
        return self.cur[0][: self.n, 1]

    @property
    def sediment(self):
        return self.cur[1][: self.n]


def _erode_terrain3_exact(nodes, neighbors, heights, num_iter, return_state):
    """float64, no FMA, reference order: bit-identical to the reference (nxb_erode3_step_f64)."""
    import ctypes as C
    from . import _lib
    adj = _neighbors(neighbors)
    if isinstance(nodes, DeviceMesh):
        n64 = rt.mesh_points(nodes.k, nodes.v_begin, nodes.v_begin + nodes.n_vertices, f32=False, f64=True,
                             device=nodes.xyz.device)[1] * nodes.radius
    elif isinstance(nodes, torch.Tensor):
        n64 = nodes
    else:
        n64 = rt.upload(np.ascontiguousarray(nodes, dtype=np.float64))
    dev_io = isinstance(heights, torch.Tensor)
    cur = [heights.clone() if dev_io else rt.upload(np.ascontiguousarray(heights, dtype=np.float64)), None, None]
    cur[1], cur[2] = torch.zeros_like(cur[0]), torch.zeros_like(cur[0])
    nxt = [torch.empty_like(t) for t in cur]
    for _ in range(num_iter):
        _lib.call("nxb_erode3_step_f64", rt._ptr(n64), rt._ptr(adj), rt._ptr(cur[0]), rt._ptr(cur[1]), rt._ptr(cur[2]),
                  rt._ptr(nxt[0]), rt._ptr(nxt[1]), rt._ptr(nxt[2]), cur[0].numel(), C.c_double(RAIN_AMOUNT), rt._stream())
        cur, nxt = nxt, cur
    if dev_io:
        return tuple(cur) if return_state else cur[0]
    heights[...] = rt._to_host(cur[0])
    if return_state:
        return rt._to_host(cur[1]), rt._to_host(cur[2])
    return None


def _snapshot(st, i):
    """erosion.py:186-192: a grayscale map of the heights after sweep i (rescale(heights, 0, 255) blended
    through the nearest-vertex query of the export path, nixis.py:270-283) saved as
    `erosion_snapshot_<iii>` in cfg.SNAP_DIR.  The heights stay on the device: one fused map kernel
    (rescale + 3-vertex blend) per snapshot."""
    import os
    from . import util
    cfg = util.cfg
    if cfg.IMG_QUERY_DATA is None:
        raise ValueError("erode_terrain3(snapshot=True) needs cfg.IMG_QUERY_DATA (nixis.py:283: the nearest-vertex "
                         "query of every pixel); run util.build_KDTree / KDT.query first")
    dists, ids = cfg.IMG_QUERY_DATA[0], cfg.IMG_QUERY_DATA[1]
    d = dists if isinstance(dists, torch.Tensor) else rt.upload(np.ascontiguousarray(dists, dtype=np.float64))
    n = ids if isinstance(ids, torch.Tensor) else rt.upload(np.ascontiguousarray(ids, dtype=np.int64))
    h = st.heights.contiguous()
    lo, hi = (float(v) for v in rt.minmax(h).tolist())
    img = rt.idw_map(d, n, h, lo, hi, 0.0, 255.0, out_bits=8)
    snap_dir = cfg.SNAP_DIR or os.getcwd()
    util.save_image({f"{i + 1:03d}": img}, snap_dir, "erosion_snapshot")


def _erode_terrain3_sharded(ctx, nodes, neighbors, heights, num_iter, return_state):
    """One rank's slice under a shard context (shard.py): halo plan from the rows given, halo positions
    fetched once for the edge lengths, then the sharded sweep loop with the fused exchange."""
    from .multigpu import ShardedErosion
    from .partition import build_rank_plan_local, exchange_halo_torch
    import os, time
    marks = [("start", time.perf_counter())]

    def mark(name):                     # NXB_TIMING=1: where the call's time goes (tools/e2e_probe.py)
        if os.environ.get("NXB_TIMING"):
            torch.cuda.synchronize()
            marks.append((name, time.perf_counter()))

    rows = _neighbors(neighbors)
    mark("upload rows")
    plan = build_rank_plan_local(rows, ctx.rank, ctx.world, ctx.ranges, group=ctx.group)
    mark("halo plan")
    own64 = nodes if isinstance(nodes, torch.Tensor) else rt.upload(np.ascontiguousarray(nodes, dtype=np.float64))
    assert own64.dtype == torch.float64 and own64.shape == (plan.n_own, 3), "nodes must be this rank's float64 [n_own,3] slice"
    comps = []
    for a in range(3):                  # positions of the halo vertices come from their owners, once
        c = torch.zeros(plan.capacity, dtype=torch.float64, device=own64.device)
        c[: plan.n_own] = own64[:, a]
        comps.append(c)
    mark("upload positions")
    exchange_halo_torch(plan, comps, group=ctx.group)
    mark("halo positions")
    dist_f32 = rt.edge_lengths(torch.stack(comps, dim=1).contiguous(), plan.local_adj)
    del comps
    mark("edge lengths")
    ero = ShardedErosion(plan, dist_f32, transport="fused", group=ctx.group)
    mark("tile plan + peer memory")
    dev_io = isinstance(heights, torch.Tensor)
    ero.load(heights if dev_io else rt.upload_f32(heights))
    mark("upload heights + publish")
    ero.run(num_iter)
    ero.finish()
    mark("sweeps")

    def report():
        if os.environ.get("NXB_TIMING"):
            print(f"  erode_terrain3 (sharded, rank {ctx.rank}): " +
                  ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.1f} ms" for a, b in zip(marks, marks[1:])), flush=True)
    try:
        if dev_io:
            return (ero.heights.clone(), ero.water.clone(), ero.sediment.clone()) if return_state else ero.heights.clone()
        rt.download_f64(ero.heights.contiguous(), out=heights)
        mark("download")
        if return_state:
            return rt.download_f64(ero.water.contiguous()), rt.download_f64(ero.sediment.contiguous())
        return None
    finally:
        ero.close()
        mark("close")
        report()


def erode_terrain3(nodes, neighbors, heights, num_iter=1, snapshot=False, verbose=True, return_state=False, exact=False):
    """erosion.py:172-192.  exact=True: float64 kernel without FMA in the reference's operation order,
    bit-identical to the reference (slower: double state, global gathers); default FP32 tile-plan kernel.
    `heights` (numpy float64) is eroded IN PLACE and None is returned; with CUDA tensors the new height
    tensor is returned.  water / sediment start at zero and are discarded unless return_state=True.
    snapshot=True saves a grayscale map after every sweep (erosion.py:186-192; needs cfg.IMG_QUERY_DATA).
    Under a shard context (shard.set_shard) the arguments are this rank's slices and the sweeps run
    sharded with the NVLink halo exchange."""
    if verbose:
        print("Starting terrain erosion...")
    if num_iter <= 0:
        num_iter = 1
    ctx = shard.current()
    if ctx is not None and ctx.world > 1:
        if exact or snapshot:
            raise ValueError("exact / snapshot modes of erode_terrain3 are single-process")
        return _erode_terrain3_sharded(ctx, nodes, neighbors, heights, num_iter, return_state)
    if exact:
        return _erode_terrain3_exact(nodes, neighbors, heights, num_iter, return_state)
    adj = _neighbors(neighbors)
    dev_io = isinstance(heights, torch.Tensor)
    h32 = heights if dev_io else rt.upload_f32(heights)
    st = Erosion3State(nodes, adj, h32)
    if snapshot:
        for i in range(num_iter):
            if verbose:
                print("  Erosion pass:", i + 1, "of", num_iter)
            st.step()
            _snapshot(st, i)
    else:
        if verbose:                     # launches are asynchronous: the progress lines are all there is to see
            for i in range(num_iter):
                print("  Erosion pass:", i + 1, "of", num_iter)
        st.run(num_iter)                # the whole loop in one C call
    if dev_io:
        return st if return_state else st.heights
    rt.download_f64(st.heights.contiguous(), out=heights)
    if return_state:
        return rt.download_f64(st.water.contiguous()), rt.download_f64(st.sediment.contiguous())
    return None


def erosion_iteration3(verts, neighbors, r_buff, wat, sed):
    """One sweep (erosion.py:197-279): r_buff, wat, sed (numpy float64) are updated in place.
    `wat` must already contain this iteration's rain, as in erode_terrain3."""
    adj = _neighbors(neighbors)
    st = Erosion3State(verts, adj, rt.upload_f32(r_buff))
    st.cur[0][: st.n, 1].copy_(rt.upload_f32(wat))
    st.cur[1][: st.n].copy_(rt.upload_f32(sed))
    st.step(rain=0.0)
    rt.download_f64(st.heights.contiguous(), out=r_buff)
    rt.download_f64(st.water.contiguous(), out=wat)
    rt.download_f64(st.sediment.contiguous(), out=sed)


def erosion_iteration1(neighbors, r_buff, w_buff):
    """erosion.py:76-99: w = r + 0.0005 * (#higher - #lower neighbours); returns w_buff."""
    adj = _neighbors(neighbors)
    if isinstance(r_buff, torch.Tensor):
        rt.erode1_step(adj, r_buff, w_buff, 0, r_buff.numel())
        return w_buff
    src = rt.upload_f32(r_buff)
    dst = torch.empty_like(src)
    rt.erode1_step(adj, src, dst, 0, src.numel())
    return rt.download_f64(dst, out=w_buff)


def erode_terrain1(nodes, neighbors, heights, num_iter=1, snapshot=None, verbose=True):
    """erosion.py:42-73: heights updated in place after every pass and returned."""
    if verbose:
        print("Starting terrain erosion...")
    if num_iter <= 0:
        num_iter = 1
    adj = _neighbors(neighbors)
    dev_io = isinstance(heights, torch.Tensor)
    a = heights if dev_io else rt.upload_f32(heights)
    b = torch.empty_like(a)
    for i in range(num_iter):
        if verbose:
            print("  Erosion pass:", i + 1, "of", num_iter)
        rt.erode1_step(adj, a, b, 0, a.numel())
        a, b = b, a
    if dev_io:
        if a.data_ptr() != heights.data_ptr():
            heights.copy_(a)
        return heights
    return rt.download_f64(a, out=heights)
