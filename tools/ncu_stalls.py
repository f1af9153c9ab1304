"""Per-instruction stall summary of an ncu report's source page (SASS view), no GPU needed.

    python tools/ncu_stalls.py <report.ncu-rep> [top_n]
Prints the total of every stall reason over the kernel and the top instructions by samples."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
kernels = out.split('"Kernel Name",')
for blk in kernels[1:2]:
    lines = blk.splitlines()
    name = lines[0]
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = {h: 0 for h in stall_cols}
    data = []
    n_inst = 0
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        smp = int(r[ix["# Samples"]] or 0)
        ex = int(r[ix["Instructions Executed"]] or 0)
        n_inst += ex
        st = {h: int(r[ix[h]] or 0) for h in stall_cols}
        for h in stall_cols:
            tot[h] += st[h]
        data.append((smp, ex, r[ix["Source"]].strip(), st))
    all_s = sum(d[0] for d in data)
    print(name[:80], "samples", all_s, "warp instructions executed", n_inst)
    print("stall totals:", ", ".join(f"{h[6:]} {100*v/all_s:.1f}%" for h, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v))
    for smp, ex, src, st in sorted(data, key=lambda d: -d[0])[:top]:
        main = ", ".join(f"{h[6:]} {v}" for h, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
        print(f"  {100*smp/all_s:5.1f}%  exec {ex:>9d}  {src[:70]:70s} {main}")
