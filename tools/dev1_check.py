import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, ctypes as C
dev = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.cuda.set_device(dev)
from nixis_b200 import runtime as rt, _lib
from nixis_b200.pipeline import TerrainPipeline
from oracle import oracle, icosphere
L = _lib.load()
sm, clk, mem, a, b = C.c_int(), C.c_int(), C.c_int64(), C.c_int(), C.c_int()
L.nxb_device_info(C.byref(sm), C.byref(clk), C.byref(mem), C.byref(a), C.byref(b))
print("torch current", torch.cuda.current_device(), "lib sees sm", sm.value, "cc", a.value, b.value)
k = 64
pipe = TerrainPipeline(k, seed=12345, n_octaves=8)
pipe.build_mesh()
pts, cells = icosphere.icosa_sphere(k)
adj = oracle.build_adjacency(cells); oracle.sort_adjacency(adj)
print("adjacency equal:", np.array_equal(pipe.adj.cpu().numpy(), adj), pipe.adj.device)
x = torch.arange(10, device="cuda", dtype=torch.int64)
print(x.sum().item())
