"""Tiny driver for ncu captures: python tools/prof_driver.py {fbm|erode3|erode3comm|erode1|assembly} [k] [launches]
(erode3comm: the exchange-capable instantiation of the sweep, no peers, on one GPU)"""
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
    elif what == "erode3comm":
        import ctypes as C
        from nixis_b200 import _lib
        tp = pipe._plan
        ticket = torch.zeros(4, dtype=torch.int32, device="cuda")
        a, b = st.cur, st.nxt
        d3 = tp.dist3_for(st.dist)
        for rep in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.call("nxb_erode3_run_comm_f32", rt._ptr(tp.mem), rt._ptr(tp.adj), rt._ptr(st.dist), None if d3 is None else rt._ptr(d3),
                      rt._ptr(a[0]), rt._ptr(a[1]), rt._ptr(b[0]), rt._ptr(b[1]),
                      tp.n_own, C.c_float(0.3 / 320), n, None, 0, None, None, None, None, None, 0,
                      C.c_uint32(0), rt._ptr(ticket), rt._stream())
            e1.record(); torch.cuda.synchronize()
            print("erode3comm", k, n, "sweeps:", e0.elapsed_time(e1) / n * 1e3, "us/sweep")
            e0.record(); st.run(n); e1.record(); torch.cuda.synchronize()
            print("erode3    ", k, n, "sweeps:", e0.elapsed_time(e1) / n * 1e3, "us/sweep")
    else:
        a, b = st.cur[0], torch.empty_like(st.cur[0])
        for _ in range(n):
            rt.erode1_step(pipe.adj, a, b, 0, pipe.V)
torch.cuda.synchronize()
print("done", what, k, n)
