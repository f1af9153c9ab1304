import torch, sys
sys.path.insert(0, ".")
from nixis_b200.pipeline import TerrainPipeline
pipe = TerrainPipeline(2500, seed=12345, n_octaves=8); pipe.build_mesh(with_adjacency=False)
out = torch.empty(pipe.V, dtype=torch.float32, device="cuda")
for _ in range(3): pipe.fbm(out=out)
torch.cuda.synchronize()
best = 1e9
for rep in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): pipe.fbm(out=out)
    b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b) / 10)
print("fbm d=2500 8 oct: %.3f ms  (%.1f Gvert*oct/s)" % (best, pipe.V * 8 / best / 1e6))
