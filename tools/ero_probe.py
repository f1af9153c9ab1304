"""Per-sweep time of the erosion kernel variants on one GPU (CUDA events, C-side loop).

    python tools/ero_probe.py [division] [sweeps]
Variants: one stored length per edge on/off (NXB_ERO_DIST3), programmatic dependent launch on/off
(NXB_ERO_PDL), implicit adjacency on/off (NXB_ERO_AFFINE), pipeline depth (NXB_ERO_STAGES)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nixis_b200.pipeline import TerrainPipeline

k = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
pipe = TerrainPipeline(k, seed=12345, n_octaves=8)
pipe.build_mesh()
h, _, _ = pipe.heights()
st = pipe.erosion_state(h)
p = st.plan
print(f"d={k}: V={pipe.V} tiles {p.n_tiles} irregular {p.n_irregular} affine {p.n_affine} one-length-per-edge {p.n_affine3} two-piece {p.n_two}", flush=True)
ref = None
for env in ({}, {"NXB_ERO_WAIT_HINT": "500"}, {"NXB_ERO_WAIT_HINT": "2000"}, {"NXB_ERO_WAIT_HINT": "20000"}, {"NXB_ERO_WAIT_HINT": "2000", "NXB_ERO_STAGES": "4"}, {"NXB_ERO_STAGES": "4"},
            {"NXB_ERO_TWO": "0"}, {"NXB_ERO_DIST3": "0"}, {"NXB_ERO_PDL": "0"}, {"NXB_ERO_AFFINE": "0"}, {}):
    for key in ("NXB_ERO_DIST3", "NXB_ERO_PDL", "NXB_ERO_AFFINE", "NXB_ERO_STAGES", "NXB_ERO_TWO", "NXB_ERO_WAIT_HINT"):
        os.environ.pop(key, None)
    os.environ.update(env)
    best = 1e9
    for rep in range(3):
        st = pipe.erosion_state(h)
        st.run(20)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); st.run(n); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / n)
    st = pipe.erosion_state(h)
    st.run(12)
    res = (st.heights.clone(), st.water.clone(), st.sediment.clone())
    same = "ref" if ref is None else str(all(torch.equal(x, y) for x, y in zip(ref, res)))
    ref = ref or res
    print(f"{str(env):55s} {best*1e3:8.1f} us/sweep  {pipe.V/best/1e3:9.0f} Mvert-iter/s  {60*pipe.V/best/1e6:7.0f} GB/s(60B)  bit-identical to default: {same}", flush=True)
