"""Where a sweep's cycles go, per role (needs a library built with NXB_ERO_PROFILE=1, passed through NXB_SO):
    NXB_ERO_PROFILE=1 python -m nixis_b200.build --force; cp nixis_b200/libnixis_b200.so tools/variants/libnxb_prof.so
    python -m nixis_b200.build --force; NXB_SO=tools/variants/libnxb_prof.so python tools/ero_profile.py [division]
Producer warp (lane 0): cycles decoding the descriptor / waiting for a free stage / writing the header and issuing the
bulk copies.  Consumer warp 1 (lane 0): cycles waiting for the stage's data / everything else.  Averages over the CTAs of
the LAST sweep of a run."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nixis_b200 import _lib
from nixis_b200.pipeline import TerrainPipeline
k = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
pipe = TerrainPipeline(k, seed=12345, n_octaves=8)
pipe.build_mesh()
h, _, _ = pipe.heights()
lib = C.CDLL(_lib.SO_PATH)
for env in ({}, {"NXB_ERO_STAGES": "4"}, {"NXB_ERO_TWO": "0"}):
    for key in ("NXB_ERO_STAGES", "NXB_ERO_TWO"):
        os.environ.pop(key, None)
    os.environ.update(env)
    st = pipe.erosion_state(h)
    st.run(20)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); st.run(50); b.record(); torch.cuda.synchronize()
    n = 592
    buf = np.zeros((n, 8), dtype=np.uint64)
    rc = lib.nxb_erode_profile_read(buf.ctypes.data_as(C.c_void_p), n)
    assert rc == 0, rc
    p = buf.astype(np.float64)
    tiles = p[:, 6].mean()
    print(f"{env}: {a.elapsed_time(b) / 50 * 1e3:.1f} us/sweep, {tiles:.0f} tiles per CTA; cycles per tile:")
    print(f"   producer: decode {p[:,0].mean()/tiles:7.0f}  wait-empty {p[:,1].mean()/tiles:7.0f}  header {p[:,2].mean()/tiles:7.0f}  issue {p[:,3].mean()/tiles:7.0f}  total {p[:,7].mean()/tiles:7.0f}")
    print(f"   consumer: wait-full {p[:,4].mean()/tiles:7.0f}  work {p[:,5].mean()/tiles:7.0f}", flush=True)
