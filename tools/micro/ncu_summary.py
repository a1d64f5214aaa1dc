"""Compact per-kernel summary of an `ncu --page raw --csv` export: the metrics DESIGN.md / profiles/README.md quote.

  python tools/micro/ncu_summary.py gpurun_out/r2/ncu_pcg_bs3_raw.csv > profiles/r02_..._summary.json"""
import csv
import json
import sys

KEYS = {
    "gpu__time_duration.sum": "time_us",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active": "l1tex_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_cycles_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "regs",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_B",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio": "stall_membar",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
}


def main(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units, data = rows[h], rows[h + 1], rows[h + 2:]
    out = []
    for r in data:
        if len(r) != len(hdr):
            continue
        d = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")}
        for k, name in KEYS.items():
            if k in hdr:
                v = r[hdr.index(k)]
                try:
                    v = float(v.replace(",", ""))
                    u = units[hdr.index(k)]
                    if name.endswith("_MB") and u.lower().startswith("gbyte"):
                        v *= 1000.0
                    if name.endswith("_MB") and u.lower().startswith("kbyte"):
                        v /= 1000.0
                    if name == "time_us" and u.startswith("ms"):
                        v *= 1000.0
                    if name == "time_us" and u.startswith("ns"):
                        v /= 1000.0
                except ValueError:
                    pass
                d[name] = v
        out.append(d)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
