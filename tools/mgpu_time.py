"""torchrun --nproc-per-node N tools/mgpu_time.py : where a sharded sweep's time goes.

Per rank: own vertices, tile kinds, halo / send sizes, the COMPUTE-ONLY time of a sweep on the shard
(single-GPU instantiation, no exchange) and the time of the exchanging loop; variants via env
(NXB_ERO_PDL, NXB_SKELETON_COST, NXB_HALO_WAIT)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from nixis_b200 import runtime as rt
from nixis_b200.multigpu import ShardedTerrain
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
k = int(os.environ.get("MGPU_K", "2500"))
ev = lambda: torch.cuda.Event(enable_timing=True)


def gather(vals):
    t = torch.tensor(vals, dtype=torch.float64, device="cuda")
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]


for cost in os.environ.get("MGPU_COSTS", "6").split(","):
    os.environ["NXB_SKELETON_COST"] = cost
    terr = ShardedTerrain(k, seed=12345, n_octaves=8, transport="fused")
    h, _, _ = terr.heights()
    ero = terr.erosion
    tp = ero.tile_plan
    # compute only: the single-GPU instantiation on this shard's buffers
    a, b = (ero.hw[0], ero.sed[0]), (ero.hw[1], ero.sed[1])
    ero.load(h); ero.finish(); torch.cuda.synchronize(); dist.barrier()
    rt.erode3_run(tp, ero.dist, a, b, 0.0, 20)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = ev(), ev()
    e0.record(); rt.erode3_run(tp, ero.dist, a, b, 0.0, 200); e1.record(); torch.cuda.synchronize()
    compute_us = e0.elapsed_time(e1) / 200 * 1e3
    info = gather([terr.n_own, tp.n_tiles, tp.n_irregular, tp.n_affine, tp.n_affine3, terr.plan.n_halo,
                   sum(int(v.numel()) for v in terr.plan.send_idx.values()), len(ero.send_peers), len(ero.recv_peers),
                   getattr(ero, "n_send_tiles", 0), compute_us])
    if rank == 0:
        print(f"== d={k} world={world} skeleton cost {cost}")
        for r, v in enumerate(info):
            print(f"  rank {r}: own {int(v[0])} tiles {int(v[1])} irregular {int(v[2])} affine {int(v[3])} kind3 {int(v[4])} halo {int(v[5])} "
                  f"sent {int(v[6])} peers {int(v[7])}/{int(v[8])} send-tiles {int(v[9])}  compute-only {v[10]:.1f} us/sweep", flush=True)
    # the exchange-capable instantiation, step by step: (1) no sends, no waits; (2) sends + flags, no waits
    # (results are garbage, the time is what the peer stores cost); then the real loop
    import ctypes as C
    from nixis_b200 import _lib
    d3 = tp.dist3_for(ero.dist)
    pa, pb = ero._peer_arrays[0], ero._peer_arrays[1]
    for label, n_send, n_wait in (("COMM kernel, no sends, no waits", 0, 0), ("COMM kernel, sends + flags, no waits", len(ero.send_peers), 0)):
        torch.cuda.synchronize(); dist.barrier()
        def run(n):
            _lib.call("nxb_erode3_run_comm_f32", rt._ptr(tp.mem), rt._ptr(tp.adj), rt._ptr(ero.dist), None if d3 is None else rt._ptr(d3),
                      rt._ptr(a[0]), rt._ptr(a[1]), rt._ptr(b[0]), rt._ptr(b[1]), tp.n_own, C.c_float(0.0), n,
                      rt._ptr(ero.send_entries), n_send, pa[0], pb[0], pa[1], rt._ptr(ero.flags), rt._ptr(ero.recv_ranks), n_wait,
                      C.c_uint32(1 << 20), rt._ptr(ero.ticket), rt._stream())
        run(20)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = ev(), ev()
        e0.record(); run(200); e1.record(); torch.cuda.synchronize()
        t = gather([e0.elapsed_time(e1) / 200 * 1e3])
        if rank == 0:
            print(f"  {label}: per rank {[round(x[0], 1) for x in t]} us/sweep", flush=True)
        dist.barrier()
    # compute-only once more (the GPUs run at their power cap: is a later measurement slower by itself?)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = ev(), ev()
    e0.record(); rt.erode3_run(tp, ero.dist, a, b, 0.0, 200); e1.record(); torch.cuda.synchronize()
    t = gather([e0.elapsed_time(e1) / 200 * 1e3])
    if rank == 0:
        print(f"  compute-only again, after the COMM runs: per rank {[round(x[0], 1) for x in t]} us/sweep", flush=True)
    ero.sweeps = (1 << 20) + 400          # keep the flag values monotone for the real loop below
    for env in ({}, {"NXB_ERO_WAIT_IN_SWEEP": "0"}, {"NXB_ERO_PDL": "0"}):
        for key in ("NXB_ERO_PDL", "NXB_HALO_WAIT", "NXB_ERO_WAIT_IN_SWEEP"):
            os.environ.pop(key, None)
        os.environ.update(env)
        ero.wait_mode = os.environ.get("NXB_HALO_WAIT", "kernel")
        best = 1e9
        for rep in range(3):
            ero.load(h)
            ero.run(20)
            torch.cuda.synchronize(); dist.barrier()
            e0, e1 = ev(), ev()
            e0.record(); ero.run(300); ero.finish(); e1.record(); torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            best = min(best, ms.item() / 300 * 1e3)
        if rank == 0:
            print(f"  exchanging loop {str(env):28s} {best:.1f} us/sweep (max over ranks)", flush=True)
    for key in ("NXB_ERO_PDL", "NXB_HALO_WAIT", "NXB_ERO_WAIT_IN_SWEEP"):
        os.environ.pop(key, None)
    ero.close()
    del terr, ero
    torch.cuda.empty_cache()
dist.destroy_process_group()
