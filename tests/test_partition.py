"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): partition index lists are exact, the halo
exchange delivers the owners' values, and sharded erosion (oracle arithmetic per shard + exchange)
is BIT-IDENTICAL to the single-domain run."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import icosphere
from nixis_b200 import partition as P


def _mesh(k):
    from oracle import oracle as O
    pts, cells = icosphere.icosa_sphere(k)
    adj = O.build_adjacency(cells)
    O.sort_adjacency(adj)
    return pts, adj


@pytest.mark.parametrize("k,world", [(4, 2), (12, 2), (12, 3), (20, 8), (3, 4)])
def test_rank_plans_exact(k, world):
    pts, adj = _mesh(k)
    V = len(adj)
    ranges = P.vertex_ranges(V, world)
    assert ranges[0][0] == 0 and ranges[-1][1] == V
    assert all(b % P.TILE == 0 or b == V for b, _ in ranges)
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    plans = [P.build_rank_plan(torch.from_numpy(adj), r, world, ranges) for r in range(world)]
    owner = np.zeros(V, dtype=np.int64)
    for r, (b, e) in enumerate(ranges):
        owner[b:e] = r
    for r, pl in enumerate(plans):
        b, e = ranges[r]
        # plain-python construction of the same lists
        halo = sorted({int(n) for n in adj[b:e].ravel() if n >= 0 and not (b <= n < e)})
        assert pl.halo.tolist() == halo
        assert pl.capacity == P.round_up(e - b, P.TILE) + P.round_up(len(halo), P.TILE)
        pos = {g: i for i, g in enumerate(halo)}
        loc = pl.local_adj.numpy()
        for i in range(e - b):
            for q in range(6):
                g = int(adj[b + i, q])
                exp = -1 if g < 0 else (g - b if b <= g < e else pl.n_own_pad + pos[g])
                assert loc[i, q] == exp
        for p in range(world):
            mine = [g for g in halo if owner[g] == p]
            if p != r and mine:
                off, cnt = pl.recv_slice[p]
                assert halo[off:off + cnt] == mine
                # the peer's send list is exactly these vertices, in this order
                assert (plans[p].send_idx[r].numpy().astype(np.int64) + ranges[p][0]).tolist() == mine
                assert plans[p].send_dst_offset[r] == off and plans[p].peer_n_own_pad[r] == pl.n_own_pad
            else:
                assert p not in pl.recv_slice


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, k, n_iter, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        pts, adj = _mesh(k)
        V = len(adj)
        plan = P.build_rank_plan(torch.from_numpy(adj), rank, world)
        b, e, cap = plan.begin, plan.end, plan.capacity
        rng = np.random.default_rng(5)
        h_glob = rng.normal(size=V) * 1000.0
        # (1) exchange delivers owner values
        h = torch.zeros(cap, dtype=torch.float64)
        w = torch.zeros(cap, dtype=torch.float64)
        h[: e - b] = torch.from_numpy(h_glob[b:e])
        w[: e - b] = torch.from_numpy(np.arange(b, e, dtype=np.float64))
        P.exchange_halo_torch(plan, [h, w])
        halo = plan.halo.numpy()
        ok = np.array_equal(h[plan.n_own_pad: plan.n_own_pad + len(halo)].numpy(), h_glob[halo])
        ok &= np.array_equal(w[plan.n_own_pad: plan.n_own_pad + len(halo)].numpy(), halo.astype(np.float64))
        # (2) sharded erosion with oracle arithmetic, local numbering
        verts = np.zeros((cap, 3))
        verts[: e - b] = pts[b:e]
        verts[plan.n_own_pad: plan.n_own_pad + len(halo)] = pts[halo]
        loc = plan.local_adj.numpy().copy()
        hh = h.numpy().copy()
        ww = np.zeros(cap)
        ss = np.zeros(cap)
        for _ in range(n_iter):
            ww += 0.3 / 320                    # rain on every slot, halo included (erosion.py:182-183)
            O.lib().nxo_erosion_iteration3(e - b, verts, loc, hh, ww, ss)
            th, tw = torch.from_numpy(hh), torch.from_numpy(ww)
            P.exchange_halo_torch(plan, [th, tw])
        out[rank] = (ok, b, e, hh[: e - b].copy(), ww[: e - b].copy(), ss[: e - b].copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_erosion_bit_identical_gloo(world, oracle):
    k, n_iter = 10, 7
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, k, n_iter, out), nprocs=world, join=True)
    pts, adj = _mesh(k)
    V = len(adj)
    rng = np.random.default_rng(5)
    h = rng.normal(size=V) * 1000.0
    wat, sed = np.zeros(V), np.zeros(V)
    for _ in range(n_iter):
        wat += 0.3 / 320
        oracle.erosion_iteration3(pts, adj, h, wat, sed)
    for r in range(world):
        ok, b, e, hh, ww, ss = out[r]
        assert ok
        assert np.array_equal(hh, h[b:e]) and np.array_equal(ww, wat[b:e]) and np.array_equal(ss, sed[b:e])


def _local_plan_worker(rank, world, port, k, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pts, adj = _mesh(k)
        V = len(adj)
        ranges = P.vertex_ranges(V, world, front=12 + 30 * (k - 1), front_cost=6.0)
        b, e = ranges[rank]
        ref = P.build_rank_plan(torch.from_numpy(adj), rank, world, ranges)
        got = P.build_rank_plan_local(torch.from_numpy(adj[b:e].copy()), rank, world, ranges)
        same = (torch.equal(ref.halo, got.halo) and ref.capacity == got.capacity and torch.equal(ref.local_adj, got.local_adj)
                and ref.recv_slice == got.recv_slice and ref.send_dst_offset == got.send_dst_offset
                and ref.peer_n_own_pad == got.peer_n_own_pad and sorted(ref.send_idx) == sorted(got.send_idx)
                and all(torch.equal(ref.send_idx[p], got.send_idx[p]) for p in ref.send_idx))
        out[rank] = bool(same)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k,world", [(10, 2), (20, 3), (24, 4)])
def test_local_planner_matches_global_planner_gloo(k, world):
    """build_rank_plan_local (own rows + a request exchange, no O(V) table on any rank) produces exactly
    the index lists of build_rank_plan (whole table on every rank)."""
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_local_plan_worker, args=(world, port, k, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)


@pytest.mark.parametrize("V,world,front,cost", [(62500002, 8, 74982, 6.0), (10000002, 4, 29982, 6.0), (1024002, 3, 9582, 1.0),
                                                (162, 4, 42, 6.0), (12, 3, 12, 6.0)])
def test_vertex_ranges_cost_weighted(V, world, front, cost):
    """Ranges are contiguous, tile-aligned, cover [0, V) and equalise COST when the first `front`
    vertices (the mesh skeleton) are `cost` times as expensive as the rest."""
    r = P.vertex_ranges(V, world, front=front, front_cost=cost)
    assert len(r) == world and r[0][0] == 0 and r[-1][1] == V
    assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
    assert all(b % P.TILE == 0 or b == V for b, _ in r) and all(e >= b for b, e in r)

    def weight(b, e):
        f = max(0, min(e, front) - min(b, front))
        return f * cost + (e - b - f)
    w = [weight(b, e) for b, e in r]
    if V > 100 * P.TILE * world:
        assert max(w) - min(w) <= 2 * P.TILE * cost + V % P.TILE + P.TILE * world     # equal up to tile rounding
        if cost > 1:
            assert r[0][1] - r[0][0] < r[1][1] - r[1][0]                               # rank 0 owns fewer vertices
    assert P.vertex_ranges(V, world) == P.vertex_ranges(V, world, front=0, front_cost=cost)
