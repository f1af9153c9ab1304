"""torchrun --nproc-per-node N tools/e2e_probe.py : where the time of the SHARDED numpy-API erode_terrain3 call goes
(plan from rows, halo positions, edge lengths, tile plan + peer memory, upload, sweeps, download)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from nixis_b200 import runtime as rt, erosion, shard
from nixis_b200.multigpu import ShardedTerrain
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
k = int(os.environ.get("MGPU_K", "2500"))
terr = ShardedTerrain(k, seed=12345, n_octaves=8)
h, _, _ = terr.heights()
n = terr.n_own
pts = torch.empty((n, 3), dtype=torch.float64, pin_memory=True); pts.copy_(terr.xyz64)
nbr = torch.empty((n, 6), dtype=torch.int32, pin_memory=True); nbr.copy_(rt.icosa_adj_rows(k, terr.begin, terr.end))
hh = torch.empty(n, dtype=torch.float64, pin_memory=True); hh.copy_(h.double())
torch.cuda.synchronize()
terr.erosion.close()
shard.set_shard(terr.ranges)
os.environ["NXB_TIMING"] = "1"
for it in range(3):
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    erosion.erode_terrain3(pts.numpy(), nbr.numpy(), hh.numpy(), num_iter=int(os.environ.get("ITERS", "1000")), verbose=False)
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        print(f"call {it}: {1e3 * (time.perf_counter() - t0):.1f} ms total", flush=True)
dist.destroy_process_group()
