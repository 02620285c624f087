#!/usr/bin/env python
"""Condense an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into the handful of
numbers DESIGN.md / bench.py quote.  Usage: python scripts/ncu_summary.py rep.ncu-rep out.txt [kernel-substr]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg ", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg ",
    "sm__inst_executed_pipe_tensor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum ", "dram__bytes_write.sum ", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum ", "launch__registers_per_thread ", "launch__grid_size", "launch__block_size",
    "launch__cluster", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum ",
    "smsp__average_warps_issue_stalled", "smsp__average_warp_latency_issue_stalled",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    sub = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# condensed from {rep} (ncu --set full --clock-control none); per-launch values"]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if sub and sub not in name:
            continue
        lines.append(f"== launch id {r[0]}: {name[:60]}")
        stalls = []
        for i, h in enumerate(hdr):
            if r[i] in ("", "n/a"):
                continue
            hh = h.split(".", 2)[-1] if h.count(".") >= 2 and h.split(".")[0].isupper() else h
            if "issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                try:
                    stalls.append((float(r[i]), h))
                except ValueError:
                    pass
                continue
            if any((k.strip() in h) and (not k.endswith(" ") or h.endswith(k.strip())) for k in KEYS):
                lines.append(f"  {h} [{units[i]}] = {r[i]}")
        for v, h in sorted(stalls, reverse=True)[:6]:
            lines.append(f"  stall {h} = {v:.2f}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
