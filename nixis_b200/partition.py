"""Vertex-range partition of the planet across ranks and the halo plan of the erosion stencil.

Owner-computes over CONTIGUOUS ranges of the reference's vertex order (meshzoo order):
rank r owns global vertices [begin_r, end_r), boundaries aligned to the 256-vertex erosion tile.
The stencil radius is 1 and a sweep reads only the neighbours' height and water
(erosion.py:225-247), so per sweep each rank needs h and w of the non-owned vertices its rows
reference -- its HALO -- and nothing else.

Everything here is integer index logic on tensors (torch ops that run on CPU and on CUDA), done
once per mesh.  Two planners produce the same index lists (tests/test_partition.py checks both
against a plain-python construction):
  * build_rank_plan        -- from the WHOLE neighbour table (every rank derives every rank's halo
                              locally, no communication; O(V) memory per rank: small meshes, tests);
  * build_rank_plan_local  -- from the rank's OWN rows only: each rank finds its halo, asks the
                              owners for it (one all-gather of counts + one point-to-point exchange
                              of id lists) and learns from the requests it receives what it must
                              send.  No rank holds anything of size O(V).

Local numbering of rank r:   [0, n_own)                       own vertices, global id = begin + i
                             [n_own_pad, n_own_pad + n_halo)  halo slots, sorted by global id
so that runs of consecutive global ids stay consecutive locally (the tile planner of the sweep
kernel needs contiguous runs) and the halo slots that one peer fills form ONE contiguous range.
"""
from dataclasses import dataclass, field
from typing import Dict, List

import torch

TILE = 256


def round_up(n, m):
    return (n + m - 1) // m * m


def vertex_ranges(n_vertices: int, world: int, align: int = TILE, front: int = 0, front_cost: float = 1.0):
    """Contiguous, tile-aligned ranges of near-equal COST; the last rank takes the remainder.
    The first `front` vertices cost `front_cost` each, the rest 1: the mesh skeleton (12 corners + 30
    edge runs) sits at the front of the index space, its tiles gather their neighbours from global
    memory instead of staged runs and measure ~6x a regular tile, so rank 0 gets fewer vertices."""
    front = min(max(int(front), 0), n_vertices)
    extra = front * (front_cost - 1.0)
    per_cost = (n_vertices + extra) / world

    def vertex_at(cost):                      # smallest vertex count whose cumulative cost >= cost
        if cost <= front * front_cost:
            return cost / front_cost
        return cost - extra

    out, b = [], 0
    for r in range(world):
        e = n_vertices if r == world - 1 else min(round_up(int(-(-vertex_at((r + 1) * per_cost) // 1)), align), n_vertices)
        e = max(e, b)
        out.append((b, e))
        b = e
    return out


def halo_ids(adj_rows: torch.Tensor, begin: int, end: int) -> torch.Tensor:
    """Sorted unique global ids referenced by rows [begin, end) that are not owned (int64)."""
    flat = adj_rows.reshape(-1).to(torch.int64)
    outside = flat[(flat >= 0) & ((flat < begin) | (flat >= end))]
    return torch.unique(outside)            # sorted


@dataclass
class RankPlan:
    rank: int
    world: int
    begin: int
    end: int
    n_own: int
    n_own_pad: int
    halo: torch.Tensor                      # int64 [n_halo] global ids, sorted
    capacity: int                           # elements of every state buffer (own pad + halo pad)
    local_adj: torch.Tensor                 # int32 [n_own, 6] in local numbering, -1 kept
    # per peer p: the LOCAL own indices this rank sends (in p's halo order) ...
    send_idx: Dict[int, torch.Tensor] = field(default_factory=dict)
    # ... and the contiguous slice [offset, offset+count) of this rank's halo slots p fills
    recv_slice: Dict[int, tuple] = field(default_factory=dict)
    # where, inside peer p's halo slots, the values this rank sends land (offset in p's halo list)
    send_dst_offset: Dict[int, int] = field(default_factory=dict)
    peer_n_own_pad: Dict[int, int] = field(default_factory=dict)

    @property
    def n_halo(self):
        return int(self.halo.numel())

    @property
    def peers(self) -> List[int]:
        return sorted(set(self.send_idx) | set(self.recv_slice))


def build_rank_plan(adj_global: torch.Tensor, rank: int, world: int, ranges=None) -> RankPlan:
    """adj_global: int32 [V,6] sorted neighbour table of the WHOLE mesh (global ids, -1 pads)."""
    V = adj_global.shape[0]
    ranges = ranges or vertex_ranges(V, world)
    begin, end = ranges[rank]
    n_own = end - begin
    n_own_pad = round_up(n_own, TILE)
    rows = adj_global[begin:end]
    halo = halo_ids(rows, begin, end)
    n_halo = int(halo.numel())
    capacity = n_own_pad + round_up(n_halo, TILE)
    # local adjacency
    g = rows.to(torch.int64)
    owned = (g >= begin) & (g < end)
    slot = torch.searchsorted(halo, g.clamp(min=0)) if n_halo else torch.zeros_like(g)
    local = torch.where(owned, g - begin, slot + n_own_pad)
    local = torch.where(g < 0, torch.full_like(g, -1), local).to(torch.int32).contiguous()
    plan = RankPlan(rank, world, begin, end, n_own, n_own_pad, halo, capacity, local)
    bounds = torch.tensor([b for b, _ in ranges] + [V], dtype=torch.int64, device=halo.device)
    # what I receive: my halo list cut by owner
    if n_halo:
        cuts = torch.searchsorted(halo, bounds).tolist()
        for p in range(world):
            cnt = cuts[p + 1] - cuts[p]
            if p != rank and cnt > 0:
                plan.recv_slice[p] = (cuts[p], cnt)
    # what I send: every peer's halo (derived locally from the full table) restricted to my range
    for p in range(world):
        if p == rank:
            continue
        pb, pe = ranges[p]
        if pe <= pb:
            continue
        ph = halo_ids(adj_global[pb:pe], pb, pe)
        if ph.numel() == 0:
            continue
        lo, hi = torch.searchsorted(ph, torch.tensor([begin, end], dtype=torch.int64, device=ph.device)).tolist()
        if hi > lo:
            plan.send_idx[p] = (ph[lo:hi] - begin).to(torch.int32).contiguous()
            plan.send_dst_offset[p] = lo
            plan.peer_n_own_pad[p] = round_up(pe - pb, TILE)
    return plan


def exchange_halo_torch(plan: RankPlan, arrays, group=None):
    """Reference halo exchange with torch.distributed point-to-point ops (NCCL on GPUs, gloo on CPU):
    for every array in `arrays` (1-D, plan.capacity elements) send the peers' halo values and receive
    this rank's halo slots in place.  Used as the baseline / fallback of the NVLink put kernel and by
    the CPU tests."""
    import torch.distributed as dist
    ops, keep = [], []
    for p in plan.peers:
        if p in plan.send_idx:
            idx = plan.send_idx[p].to(torch.int64)
            buf = torch.stack([a[idx] for a in arrays]).contiguous()
            keep.append(buf)
            ops.append(dist.P2POp(dist.isend, buf, p, group=group))
        if p in plan.recv_slice:
            off, cnt = plan.recv_slice[p]
            rbuf = torch.empty((len(arrays), cnt), dtype=arrays[0].dtype, device=arrays[0].device)
            keep.append((rbuf, off, cnt))
            ops.append(dist.P2POp(dist.irecv, rbuf, p, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for item in keep:
        if isinstance(item, tuple):
            rbuf, off, cnt = item
            for i, a in enumerate(arrays):
                a[plan.n_own_pad + off: plan.n_own_pad + off + cnt] = rbuf[i]


def _p2p_exchange(send: Dict[int, torch.Tensor], recv_counts: Dict[int, int], dtype, device, group=None):
    """Variable-size point-to-point exchange: send[p] goes to rank p, recv_counts[p] elements come
    back from rank p.  gloo moves CPU tensors, NCCL device tensors."""
    import torch.distributed as dist
    cpu = dist.get_backend(group) == "gloo"
    ops, keep, out = [], [], {}
    for p, t in send.items():
        buf = (t.cpu() if cpu else t).contiguous()
        keep.append(buf)
        ops.append(dist.P2POp(dist.isend, buf, p, group=group))
    for p, cnt in recv_counts.items():
        buf = torch.empty(cnt, dtype=dtype, device="cpu" if cpu else device)
        out[p] = buf
        ops.append(dist.P2POp(dist.irecv, buf, p, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return {p: b.to(device) for p, b in out.items()}


def build_rank_plan_local(rows: torch.Tensor, rank: int, world: int, ranges, group=None) -> RankPlan:
    """rows: int32 [n_own, 6] sorted neighbour rows of THIS rank's range (global ids, -1 pads).
    Collective over `group`: every rank calls it with its own rows."""
    import torch.distributed as dist
    begin, end = ranges[rank]
    n_own = end - begin
    assert rows.shape[0] == n_own
    dev = rows.device
    n_own_pad = round_up(n_own, TILE)
    halo = halo_ids(rows, begin, end)
    n_halo = int(halo.numel())
    capacity = n_own_pad + round_up(n_halo, TILE)
    g = rows.to(torch.int64)
    owned = (g >= begin) & (g < end)
    slot = torch.searchsorted(halo, g.clamp(min=0)) if n_halo else torch.zeros_like(g)
    local = torch.where(owned, g - begin, slot + n_own_pad)
    local = torch.where(g < 0, torch.full_like(g, -1), local).to(torch.int32).contiguous()
    plan = RankPlan(rank, world, begin, end, n_own, n_own_pad, halo, capacity, local)
    V = ranges[-1][1]
    bounds = torch.tensor([b for b, _ in ranges] + [V], dtype=torch.int64, device=dev)
    cuts = torch.searchsorted(halo, bounds).tolist() if n_halo else [0] * (world + 1)
    want = [cuts[p + 1] - cuts[p] if p != rank else 0 for p in range(world)]      # ids I need from rank p
    for p in range(world):
        if want[p] > 0:
            plan.recv_slice[p] = (cuts[p], want[p])
    if world == 1:
        return plan
    # who asks whom for how much: counts[q][p] = ids rank q needs from rank p
    cpu = dist.get_backend(group) == "gloo"
    mine = torch.tensor(want, dtype=torch.int64, device="cpu" if cpu else dev)
    allc = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine, group=group)
    counts = torch.stack(allc).cpu()
    requests = _p2p_exchange({p: halo[cuts[p]:cuts[p + 1]] for p in range(world) if want[p] > 0},
                             {q: int(counts[q, rank]) for q in range(world) if q != rank and int(counts[q, rank]) > 0},
                             torch.int64, dev, group)
    for q, ids in requests.items():
        plan.send_idx[q] = (ids - begin).to(torch.int32).contiguous()
        plan.send_dst_offset[q] = int(counts[q, :rank].sum())       # my block inside q's (sorted) halo list
        plan.peer_n_own_pad[q] = round_up(ranges[q][1] - ranges[q][0], TILE)
    return plan
