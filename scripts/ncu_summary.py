#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page raw --csv` output to the columns the roofline discussion uses (one row per launch).
    python scripts/ncu_summary.py gpurun_out/x.raw.csv profiles/rNN_x.summary.csv [max_rows]
Also prints a one-line digest per kernel name (mean duration, instructions, DRAM bytes)."""
import csv
import sys

WANT = ("gpu__time_duration", "launch__registers", "launch__shared_mem", "launch__occupancy_limit", "launch__waves",
        "launch__grid_size", "launch__block_size", "sm__warps_active", "smsp__inst_executed.sum", "smsp__issue_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate", "l1tex__t_sector_hit_rate",
        "l1tex__throughput", "lts__throughput", "sm__throughput", "dram__throughput", "issue_stalled", "bank_conflicts",
        "thread_inst_executed_per_inst", "sm__cycles_active.avg", "sm__inst_executed_pipe", "smsp__inst_executed_pipe",
        "sm__pipe_fma", "sm__pipe_alu", "sm__pipe_xu", "sm__pipe_tensor")


def main():
    src, dst = sys.argv[1], sys.argv[2]
    max_rows = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr, data = rows[hi], rows[hi + 2:]
    keep = [i for i, h in enumerate(hdr) if h in ("ID", "Kernel Name") or any(s in h for s in WANT)]
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        for r in rows[hi:hi + 2] + data[:max_rows]:
            w.writerow([r[i] for i in keep])
    col = {h: i for i, h in enumerate(hdr)}
    by = {}
    for r in data:
        by.setdefault(r[col["Kernel Name"]].split("(")[0], []).append(r)

    def mean(rs, name):
        return sum(float(r[col[name]].replace(",", "")) for r in rs) / len(rs) if name in col else float("nan")

    for k, rs in by.items():
        unit = rows[hi + 1][col["gpu__time_duration.sum"]]
        print(f"{k[:70]:70s} n={len(rs)} dur={mean(rs, 'gpu__time_duration.sum'):.4g} {unit} "
              f"inst={mean(rs, 'smsp__inst_executed.sum'):.4g} dram_rd={mean(rs, 'dram__bytes_read.sum'):.4g} "
              f"{rows[hi + 1][col['dram__bytes_read.sum']]} dram_wr={mean(rs, 'dram__bytes_write.sum'):.4g} "
              f"warps_active={mean(rs, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.3g}% "
              f"sm_thr={mean(rs, 'sm__throughput.avg.pct_of_peak_sustained_elapsed'):.3g}%")


if __name__ == "__main__":
    main()
