"""Sustained sweep time (power-capped regime): N sweeps back to back at d=2500, clocks sampled."""
import os, sys, subprocess, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from nixis_b200.pipeline import TerrainPipeline
n = int(sys.argv[1]) if len(sys.argv) > 1 else 800
pipe = TerrainPipeline(2500, seed=12345, n_octaves=8)
pipe.build_mesh()
h, _, _ = pipe.heights()
st = pipe.erosion_state(h)
st.run(50); torch.cuda.synchronize()
clk = []
stop = False
def pump():
    while not stop:
        o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip().split(",")
        try: clk.append((float(o[0]), float(o[1])))
        except Exception: pass
        time.sleep(0.05)
t = threading.Thread(target=pump); t.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); st.run(n); e1.record(); torch.cuda.synchronize()
stop = True; t.join()
ms = e0.elapsed_time(e1)
mid = clk[len(clk)//4:] or [(0, 0)]
print(f"stages={os.environ.get('NXB_ERO_STAGES','3')}: {ms/n*1e3:.1f} us/sweep over {n} sweeps; "
      f"SM {sorted(c for c,_ in mid)[len(mid)//2]:.0f} MHz, {max(p for _,p in mid):.0f} W")
