"""BASELINE configs[4] sizing check on ONE GPU: d=5000 (250 000 002 vertices), 12-octave 4-D fBm + erosion."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nixis_b200 import runtime as rt
from nixis_b200.pipeline import TerrainPipeline, assemble_heights
k = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
t0 = time.perf_counter()
pipe = TerrainPipeline(k, seed=12345, n_octaves=12)
pipe.build_mesh()
torch.cuda.synchronize()
print(f"d={k}: V={pipe.V} mesh+adjacency {time.perf_counter()-t0:.2f}s, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
ev = lambda: torch.cuda.Event(enable_timing=True)
e0, e1, e2, e3 = ev(), ev(), ev(), ev()
w = [0.5 * f for f in pipe.freq]
e0.record()
h4 = rt.fbm4(pipe.tables, pipe.mesh.xyz, pipe.freq, pipe.amp, w)
e1.record()
h3 = pipe.fbm()
e2.record()
torch.cuda.synchronize()
print(f"fbm4 12 oct: {e0.elapsed_time(e1):.1f} ms -> {pipe.V*12/e0.elapsed_time(e1)/1e3:.0f} Mvert-oct/s; fbm3 12 oct: {e1.elapsed_time(e2):.1f} ms -> {pipe.V*12/e1.elapsed_time(e2)/1e3:.0f} Mvert-oct/s", flush=True)
h, _, lvl = assemble_heights(h4)
del h3
st = pipe.erosion_state(h)
torch.cuda.synchronize()
print(f"erosion plan: tiles {st.plan.n_tiles} irregular {st.plan.n_irregular}; mem now {torch.cuda.memory_allocated()/2**30:.1f} GiB peak {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
st.run(3)
e2.record(); st.run(20); e3.record(); torch.cuda.synchronize()
ms = e2.elapsed_time(e3) / 20
print(f"erode3: {ms:.3f} ms/sweep -> {pipe.V/ms/1e3:.0f} Mvert-iter/s, {60*pipe.V/ms/1e6:.0f} GB/s(alg); finite heights: {bool(torch.isfinite(st.heights).all())}", flush=True)
