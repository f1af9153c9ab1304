"""Turn ncu reports brought back from the GPU box into the committed evidence under profiles/ (no GPU needed).

    python tools/ncu_summary.py <key> <report.ncu-rep> [<key> <report> ...]   ->  profiles/r02_ncu_summary.json
                                                                               profiles/r02_<key>_raw.csv      (every metric ncu holds, --page raw)
                                                                               profiles/r02_<key>_stalls.txt   (per-instruction stall samples, SASS)
The JSON holds, per key and per captured launch, the metrics DESIGN.md and bench.py quote (same names as ncu's)."""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "smsp__inst_executed_op_tma_ld.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def to_bytes(val, unit):
    return float(val) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(unit, 1.0)


def main():
    args = sys.argv[1:]
    out_path = os.path.join(ROOT, "profiles", "r02_ncu_summary.json")
    summary = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for key, rep in zip(args[0::2], args[1::2]):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        open(os.path.join(ROOT, "profiles", f"r02_{key}_raw.csv"), "w").write(raw)
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        launches = []
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            rec = {"kernel": d.get("Kernel Name", "")}
            for w in WANT:
                if w in d and d[w] != "":
                    rec[w] = f"{d[w]} {units[hdr.index(w)]}".strip()
            try:
                rec["dram_bytes_per_launch"] = (to_bytes(d["dram__bytes_read.sum"], units[hdr.index("dram__bytes_read.sum")])
                                                + to_bytes(d["dram__bytes_write.sum"], units[hdr.index("dram__bytes_write.sum")]))
            except Exception:
                pass
            launches.append(rec)
        entry = dict(launches[-1]) if launches else {}
        entry["launches_captured"] = len(launches)
        entry["all_launches"] = launches
        entry["report"] = os.path.basename(rep)
        entry["raw_csv"] = f"profiles/r02_{key}_raw.csv"
        summary[key] = entry
        st = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_stalls.py"), rep, "40"], capture_output=True, text=True).stdout
        open(os.path.join(ROOT, "profiles", f"r02_{key}_stalls.txt"), "w").write(st)
        print(key, {k: entry.get(k) for k in ("gpu__time_duration.sum", "dram_bytes_per_launch", "smsp__issue_active.avg.pct_of_peak_sustained_active")})
    json.dump(summary, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
