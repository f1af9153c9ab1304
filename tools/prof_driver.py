"""Tiny driver for ncu captures: python tools/prof_driver.py {fbm|erode3|erode1|assembly} [k] [launches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nixis_b200 import runtime as rt
from nixis_b200.pipeline import TerrainPipeline, assemble_heights

what = sys.argv[1] if len(sys.argv) > 1 else "erode3"
k = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
n = int(sys.argv[3]) if len(sys.argv) > 3 else 6
pipe = TerrainPipeline(k, seed=12345, n_octaves=8)
pipe.build_mesh(with_adjacency=(what != "fbm"))
if what == "fbm":
    out = torch.empty(pipe.V, dtype=torch.float32, device="cuda")
    for _ in range(n):
        pipe.fbm(out=out)
elif what == "assembly":
    for _ in range(n):
        pipe.heights()
else:
    h, _, _ = pipe.heights()
    st = pipe.erosion_state(h.clone())
    if what == "erode3":
        st.run(n)
    else:
        a, b = st.cur[0], torch.empty_like(st.cur[0])
        for _ in range(n):
            rt.erode1_step(pipe.adj, a, b, 0, pipe.V)
torch.cuda.synchronize()
print("done", what, k, n)
