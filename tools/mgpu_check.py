"""torchrun --nproc-per-node N tools/mgpu_check.py : sharded == single-GPU (bit-identical) + timings.

MGPU_SAME_DEVICE=1: every rank uses cuda:0 (process group on gloo, peer memory through CUDA IPC): the
fused exchange -- peer stores, flags, flag waits, the C-side sweep loop -- runs for real on a one-GPU box."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from nixis_b200 import runtime as rt
from nixis_b200.multigpu import ShardedTerrain
from nixis_b200.pipeline import TerrainPipeline

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
SAME = os.environ.get("MGPU_SAME_DEVICE") == "1"
if SAME:
    torch.cuda.set_device(0)
    dist.init_process_group("gloo")
else:
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def gather_own(t, terr):
    """own slices of every rank concatenated in rank order (all_gather with unequal sizes: pad)"""
    sizes = [e - b for b, e in terr.ranges]
    m = max(sizes)
    pad = torch.zeros(m, dtype=t.dtype, device=t.device); pad[: t.numel()] = t
    if SAME:
        pad = pad.cpu()
    bufs = [torch.empty_like(pad) for _ in sizes]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)]).to(t.device)


TRANSPORTS = os.environ.get("MGPU_TRANSPORTS", "fused,nvlink" if SAME else "fused,nvlink,p2p").split(",")
SIZES = ((64, 25), (200, 10)) + (((int(os.environ["MGPU_CHECK_K"]), 6),) if os.environ.get("MGPU_CHECK_K") else ())
for transport in TRANSPORTS:
    for k, sweeps in SIZES:
        terr = ShardedTerrain(k, seed=12345, n_octaves=8, transport=transport)
        assert terr.erosion.transport == transport, (terr.erosion.transport, transport)
        h, ocean, lvl = terr.heights()
        terr.erosion.load(h)
        terr.erosion.run(sweeps - 3)            # the C-side loop ...
        for _ in range(3):
            terr.erosion.step()                 # ... and single sweeps, both parities
        terr.erosion.finish()
        torch.cuda.synchronize()
        H = gather_own(terr.erosion.heights.contiguous(), terr)
        W = gather_own(terr.erosion.water.contiguous(), terr)
        S = gather_own(terr.erosion.sediment.contiguous(), terr)
        if rank == 0:
            pipe = TerrainPipeline(k, seed=12345, n_octaves=8)
            pipe.build_mesh()
            h1, _, lvl1 = pipe.heights()
            st = pipe.erosion_state(h1)
            st.run(sweeps)
            torch.cuda.synchronize()
            same = (torch.equal(H, st.heights), torch.equal(W, st.water), torch.equal(S, st.sediment))
            print(f"[{transport}] k={k} world={world} sweeps={sweeps}: bit-identical h/w/s = {same}  level {lvl} vs {lvl1}  "
                  f"halo {terr.plan.n_halo} irregular tiles {terr.erosion.tile_plan.n_irregular}/{terr.erosion.tile_plan.n_tiles} "
                  f"peer memory {terr.erosion.peer_mem.kind if terr.erosion.peer_mem else None}", flush=True)
            assert all(same), "sharded result differs from single GPU"
        dist.barrier()
        terr.erosion.close()
        del terr
        torch.cuda.empty_cache()

# timings at scale
k = int(os.environ.get("MGPU_K", "2500"))
for transport in ([] if os.environ.get("MGPU_SKIP_TIMING") else TRANSPORTS):
    terr = ShardedTerrain(k, seed=12345, n_octaves=8, transport=transport)
    h, _, _ = terr.heights()
    ero = terr.erosion
    for n in (50, 300):
        ero.load(h)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ero.run(n); ero.finish(); e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cpu" if SAME else "cuda"); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f"[{transport}] k={k} world={world}: {n} sweeps {ms.item():.2f} ms -> {ms.item()/n*1e3:.1f} us/sweep, "
                  f"{terr.V*n/ms.item()/1e3:.0f} Mvert-iter/s; n_own {terr.n_own} halo {terr.plan.n_halo} "
                  f"irregular {ero.tile_plan.n_irregular}/{ero.tile_plan.n_tiles} setup {terr.setup_ms}", flush=True)
    t0 = time.perf_counter(); hh, _, _ = terr.heights(); torch.cuda.synchronize(); dist.barrier()
    if rank == 0: print(f"[{transport}] fbm+assembly {1e3*(time.perf_counter()-t0):.2f} ms", flush=True)
    ero.close()
    del terr, ero
    torch.cuda.empty_cache()
dist.destroy_process_group()
