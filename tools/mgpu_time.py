import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from nixis_b200.multigpu import ShardedTerrain
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
k = int(os.environ.get("MGPU_K", "2500"))
for transport in os.environ.get("MGPU_TRANSPORTS", "fused,nvlink").split(","):
    terr = ShardedTerrain(k, seed=12345, n_octaves=8, transport=transport)
    h, _, _ = terr.heights()
    ero = terr.erosion
    for n in (100, 400):
        ero.load(h)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ero.run(n); ero.finish(); e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"[{transport}{' sepwait' if os.environ.get('NXB_FUSED_SEPARATE_WAIT') else ''}] k={k} world={world}: {n} sweeps {ms.item():.2f} ms -> {ms.item()/n*1e3:.1f} us/sweep", flush=True)
    # per-rank compute-only time (no exchange): load imbalance
    from nixis_b200 import runtime as rt
    src = ero.hw[0] + (ero.sed[0],); dst = ero.hw[1] + (ero.sed[1],)
    for _ in range(5): rt.erode3_step(ero.tile_plan, ero.dist, src, dst, 0.0)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): rt.erode3_step(ero.tile_plan, ero.dist, src, dst, 0.0); src, dst = dst, src
    e1.record(); torch.cuda.synchronize()
    mine = torch.tensor([e0.elapsed_time(e1) * 10.0], device="cuda")
    allt = [torch.zeros(1, device="cuda") for _ in range(world)]
    dist.all_gather(allt, mine)
    irr = torch.tensor([float(ero.tile_plan.n_irregular)], device="cuda"); alli = [torch.zeros(1, device="cuda") for _ in range(world)]
    dist.all_gather(alli, irr)
    if rank == 0:
        print(f"[{transport}] compute-only us/sweep per rank: {[round(t.item(), 1) for t in allt]} irregular tiles per rank: {[int(t.item()) for t in alli]} of {ero.tile_plan.n_tiles}", flush=True)
    tk = ero.ticket.tolist()
    rows = sorted([tk[4 + 8 * i: 4 + 8 * i + 8] for i in range(32)], key=lambda r: r[0] & 0xffffffff)
    rows = [[x & 0xffffffff for x in r] for r in rows if r[0]]
    base = rows[0][0] if rows else 0
    us = lambda r, i: f"{(r[i] - r[0]) / 1e3:.0f}" if r[i] else "-"
    print(f"rank {rank} timeline (start us | +producers dry, +consumers dry, +last CTA past fence, +flag raised): " +
          " | ".join(f"{(r[0]-base)/1e3:.0f}: +{us(r,4)} +{us(r,5)} +{us(r,6)} +{us(r,2)}" for r in rows[:8]), flush=True)
    print(f"rank {rank} [{transport}] debug ticket words {tk[:4]} halo tiles {getattr(ero, 'n_halo_tiles', None)} boundary tiles {getattr(ero, 'n_boundary_tiles', None)} mode {getattr(ero, 'fused_mode', None)} of {ero.tile_plan.n_tiles}", flush=True)
    del terr, ero
    torch.cuda.empty_cache()
dist.destroy_process_group()
