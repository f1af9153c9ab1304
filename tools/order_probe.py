"""Single-GPU probe of the sweep kernel: implicit adjacency on / off, single-GPU vs exchange-capable
instantiation (run(order) can also time an explicit tile order)."""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nixis_b200 import runtime as rt, _lib
from nixis_b200.pipeline import TerrainPipeline
k = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
pipe = TerrainPipeline(k, seed=12345, n_octaves=8)
pipe.build_mesh()
h, _, _ = pipe.heights()
st = pipe.erosion_state(h)
tp, n_tiles = st.plan, st.plan.n_tiles
ticket = torch.zeros(4 + 128, dtype=torch.int32, device="cuda")
def run(order, n=50):
    src, dst = st.cur, st.nxt
    def one():
        nonlocal src, dst
        _lib.call("nxb_erode3_plan_step_comm_f32", rt._ptr(tp.mem), rt._ptr(tp.adj), rt._ptr(st.dist), None,
                  rt._ptr(src[0]), rt._ptr(src[1]), rt._ptr(src[2]), rt._ptr(dst[0]), rt._ptr(dst[1]), rt._ptr(dst[2]),
                  tp.n_own, C.c_float(0.0), None, None, 0, None, None, None, None, None, 0,
                  C.c_uint32(0), C.c_uint32(0), 0, rt._ptr(ticket), None if order is None else rt._ptr(order), 0, rt._stream())
        src, dst = dst, src
    for _ in range(5): one()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): one()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
ident = torch.arange(n_tiles, dtype=torch.int32, device="cuda")
print(f"k={k} tiles={n_tiles} affine={tp.n_affine} irregular={tp.n_irregular}")
def run_plain(n=50):
    src, dst = st.cur, st.nxt
    for _ in range(5):
        rt.erode3_step(tp, st.dist, src, dst, 0.0); src, dst = dst, src
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        rt.erode3_step(tp, st.dist, src, dst, 0.0); src, dst = dst, src
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for aff in ("1", "0", "1"):
    os.environ["NXB_ERO_AFFINE"] = aff
    print(f"NXB_ERO_AFFINE={aff}: single-GPU kernel {run_plain():.1f} us, exchange-capable kernel {run(None):.1f} us")
