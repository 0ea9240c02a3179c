"""Per-launch extract of an `ncu --set full` report: the columns profiles/r1_ncu_full_summary.csv keeps, plus the DRAM
traffic of the EM launches as JSON (feeds bench.py's roofline.traffic).
    ncu -i report.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv profiles/r2_ncu_full_summary.csv profiles/r2_traffic.json cfg4_chr14
The JSON is keyed by workload; bench.py uses an entry only for that workload and only while its launch count still matches.
"""
import csv
import json
import sys

KEEP = [
    "ID", "Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__cluster_dim_x",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]


def to_bytes(v, unit):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return 0.0
    return 0.0 if x != x else x * mult          # a launch with no units to run reports n/a


def main(raw, out_csv, out_json, workload="cfg4_chr14"):
    rows = list(csv.reader(open(raw)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    cols = [names.index(k) for k in KEEP if k in names]
    with open(out_csv, "w", newline="") as f:
        w = csv.writer(f, quoting=csv.QUOTE_ALL)
        w.writerow([names[c] for c in cols])
        w.writerow([units[c] for c in cols])
        for r in rows[hdr + 2:]:
            if len(r) == len(names):
                w.writerow([r[c] for c in cols])
    ki, ri, wi = names.index("Kernel Name"), names.index("dram__bytes_read.sum"), names.index("dram__bytes_write.sum")
    ti = names.index("gpu__time_duration.sum")
    rd = wr = 0.0
    n = n_ok = 0
    t_all = t_ok = 0.0
    for r in rows[hdr + 2:]:
        if len(r) == len(names) and "fb_em_kernel" in r[ki]:
            n += 1
            t = float(r[ti].replace(",", ""))
            t_all += t
            try:
                ok = float(r[ri].replace(",", "")) == float(r[ri].replace(",", ""))      # not NaN
            except ValueError:
                ok = False
            if ok:                      # ncu cuts the replay of some long launches short: they come back with their time only
                rd += to_bytes(r[ri], units[ri])
                wr += to_bytes(r[wi], units[wi])
                n_ok += 1
                t_ok += t
    scale = t_all / t_ok if t_ok else float("nan")   # the launches without counters are filled in by their share of device time
    import os
    doc = json.load(open(out_json)) if os.path.exists(out_json) else {}
    doc[workload] = {"source": f"{out_csv} (ncu --set full --profile-from-start off over the launches of the one timed step of FB_PROFILE_STEP=1 bench.py --config {workload} --steps 1 --warmup 3)",
                     "kernel": "fb_em_kernel", "launches_per_step": n, "launches_with_counters": n_ok,
                     "device_time_share_with_counters": t_ok / t_all if t_all else None,
                     "dram_bytes_read_measured": rd, "dram_bytes_write_measured": wr,
                     "dram_bytes_per_step_scaled": (rd + wr) * scale,
                     "dram_bytes_per_launch": (rd + wr) * scale / max(n, 1)}
    json.dump(doc, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:5])
