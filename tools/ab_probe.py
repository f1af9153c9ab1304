"""Same-box A/B of kernel variants: python tools/ab_probe.py [division] [sweeps] [rounds]
Every variant library under tools/variants/ (libnxb_<tag>.so, built from other commits) and the in-tree library are
measured in fresh subprocesses, round-robin, so that box-to-box and thermal drift (+-4 % on this pool) cancel."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT)
    import torch
    from nixis_b200.pipeline import TerrainPipeline
    k, n = int(sys.argv[2]), int(sys.argv[3])
    pipe = TerrainPipeline(k, seed=12345, n_octaves=8)
    pipe.build_mesh()
    h, _, _ = pipe.heights()
    best = 1e9
    for rep in range(3):
        st = pipe.erosion_state(h)
        st.run(20)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); st.run(n); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / n)
    print(f"{best * 1e3:.1f}")
    sys.exit(0)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 2500
n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
variants = {os.path.basename(p)[7:-3]: p for p in sorted(glob.glob(os.path.join(ROOT, "tools", "variants", "libnxb_*.so")))}
variants["tree"] = os.path.join(ROOT, "nixis_b200", "libnixis_b200.so")
res = {t: [] for t in variants}
for r in range(rounds):
    for tag, path in variants.items():
        out = subprocess.run([sys.executable, __file__, "--one", str(k), str(n)], env=dict(os.environ, NXB_SO=path),
                             capture_output=True, text=True)
        try:
            res[tag].append(float(out.stdout.strip().splitlines()[-1]))
        except Exception:
            res[tag].append(float("nan")); print(tag, "failed:", out.stderr[-500:])
    print(f"round {r}: " + "  ".join(f"{t} {v[-1]:.1f}" for t, v in res.items()), flush=True)
for t, v in res.items():
    print(f"{t:8s} min {min(v):.1f}  median {sorted(v)[len(v) // 2]:.1f} us/sweep")
