"""Scratch measurements on the GPU box (not part of the product): error distributions and kernel timings."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from nixis_b200 import runtime as rt, util, terrain, opensimplex as osi
from nixis_b200.pipeline import TerrainPipeline, assemble_heights
from oracle import oracle, icosphere

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts)/len(ts)

print("device", torch.cuda.get_device_name(0))
print("ffma peak TF/s", rt.ffma_peak_tflops())
# ---- error distribution fBm k=320
k = 320
mesh = util.create_mesh(k, device=True, verbose=False)
perm, pgi = osi.init(12345)
pts = mesh.points_numpy()
for n_oct in (1, 8, 12):
    h = terrain.sample_octaves(mesh, None, perm, pgi, n_oct, 1.5, 0.4, 2.5, 0.5, 1.0, verbose=False).cpu().numpy().astype(np.float64)
    ref = oracle.sample_octaves(pts, None, perm, pgi, n_oct, 1.5, 0.4, 2.5, 0.5, 1.0)
    err = np.abs(h - ref) / (ref.max() - ref.min())
    print(f"fbm k={k} oct={n_oct}: max {err.max():.3e} p99.99 {np.quantile(err,0.9999):.3e} p99 {np.quantile(err,0.99):.3e} mean {err.mean():.3e} n>1e-5 {(err>1e-5).sum()}")
# single octaves at high frequency
for f in (1.5, 58.6, 915.5, 35763.0):
    tables = rt.tables_for(perm, pgi)
    h = rt.fbm3(tables, mesh.xyz, [f], [2.0]).cpu().numpy().astype(np.float64) - 1.0
    ref = oracle.noisearr3d(pts[:,0]*f, pts[:,1]*f, pts[:,2]*f, perm, pgi)
    err = np.abs(h - ref)
    print(f"noise f={f}: max abs {err.max():.3e} p99.99 {np.quantile(err,0.9999):.3e} mean {err.mean():.3e}")
# ---- timings
for k in (1000, 2500):
    pipe = TerrainPipeline(k, seed=12345, n_octaves=8)
    t0 = time.perf_counter(); pipe.build_mesh(); torch.cuda.synchronize(); t_mesh = time.perf_counter() - t0
    V = pipe.V
    out = torch.empty(V, dtype=torch.float32, device="cuda")
    for n_oct in (1, 8):
        pipe.freq, pipe.amp = rt.octave_schedule(n_oct, 1.5, 0.4, 2.5, 0.5)
        best, avg = timeit(lambda: pipe.fbm(out=out))
        print(f"k={k} fbm {n_oct} oct: best {best:.3f} ms avg {avg:.3f} ms -> {V*n_oct/best/1e3:.1f} Mvert-oct/s, {152.4*V*n_oct/best/1e9:.2f} TFLOP/s(alg)")
    # per-octave cost at each frequency
    for o in range(8):
        f = 1.5 * 2.5**o
        pipe.freq, pipe.amp = [f], [0.4]
        best, avg = timeit(lambda: pipe.fbm(out=out), n=3, warm=1)
        print(f"   octave {o} f={f:.1f}: {best:.3f} ms -> {V/best/1e3:.1f} Mvert-oct/s")
    pipe.freq, pipe.amp = rt.octave_schedule(8, 1.5, 0.4, 2.5, 0.5)
    h, _, _ = pipe.heights()
    best, avg = timeit(lambda: assemble_heights(pipe.fbm(out=out)), n=3, warm=1)
    print(f"k={k} fbm+assembly: {best:.3f} ms; mesh+adjacency build {t_mesh*1e3:.1f} ms")
    t0 = time.perf_counter(); st = pipe.erosion_state(h.clone()); torch.cuda.synchronize()
    print(f"k={k} erosion plan: {time.perf_counter()-t0:.3f}s tiles {st.plan.n_tiles} irregular {st.plan.n_irregular} max_halo {st.plan.max_halo}")
    best, avg = timeit(lambda: st.step(), n=20, warm=3)
    print(f"k={k} erode3 step: best {best:.4f} ms avg {avg:.4f} -> {V/best/1e3:.1f} Mvert-iter/s, {60*V/best/1e6:.1f} GB/s(alg)")
    a = st.cur[0]; b = torch.empty_like(a)
    best, avg = timeit(lambda: rt.erode1_step(pipe.adj, a, b, 0, V), n=20, warm=3)
    print(f"k={k} erode1 step: best {best:.4f} ms -> {V/best/1e3:.1f} Mvert-iter/s, {32*V/best/1e6:.1f} GB/s(alg)")
    best, avg = timeit(lambda: b.copy_(a), n=20, warm=3)
    print(f"k={k} torch copy {V*4/1e6:.0f} MB: {best:.4f} ms -> {8*V/best/1e6:.1f} GB/s")
    del pipe, st, out, h, a, b
    torch.cuda.empty_cache()
