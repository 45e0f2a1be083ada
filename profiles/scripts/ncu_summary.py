#!/usr/bin/env python
"""Summarise `ncu --page raw --csv` exports: one JSON object per capture with the metrics DESIGN.md quotes.
usage: python profiles/scripts/ncu_summary.py gpurun_out/*_raw.csv > profiles/<name>.json"""
import csv
import json
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__inst_executed.sum": "warp_insts",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__inst_issued.avg.pct_of_peak_sustained_active": "inst_issued_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_wavefront_pct",
    "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed": "lsu_writeback_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__occupancy_limit_registers": "occ_limit_regs",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard": "stall_long_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard": "stall_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_wait": "stall_wait",
    "smsp__pcsamp_warps_issue_stalled_not_selected": "stall_not_selected",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle": "stall_math_pipe",
    "smsp__pcsamp_warps_issue_stalled_lg_throttle": "stall_lg_throttle",
    "smsp__pcsamp_warps_issue_stalled_mio_throttle": "stall_mio_throttle",
    "smsp__pcsamp_warps_issue_stalled_barrier": "stall_barrier",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving": "stall_branch",
    "smsp__pcsamp_warps_issue_stalled_selected": "stall_selected",
    "smsp__pcsamp_sample_buffer_full": None,
}


def summarise(path):
    rows = list(csv.reader(open(path, newline="")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    out = []
    for r in rows[hdr + 2:]:
        if len(r) != len(names):
            continue
        d = {"kernel": r[names.index("Kernel Name")]}
        for k, short in KEYS.items():
            if short and k in names:
                v = r[names.index(k)].replace(",", "")
                try:
                    d[short] = float(v)
                except ValueError:
                    d[short] = v
                d[short + "_unit"] = units[names.index(k)]
        stalls = {k: v for k, v in d.items() if k.startswith("stall_") and not k.endswith("_unit")}
        tot = sum(stalls.values()) or 1.0
        d["stall_share"] = {k[6:]: round(v / tot, 3) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1]) if v / tot >= 0.02}
        for k in list(d):
            if k.startswith("stall_") and k != "stall_share":
                del d[k]
        out.append(d)
    return out


if __name__ == "__main__":
    res = {}
    for p in sys.argv[1:]:
        res[p.split("/")[-1].replace("_raw.csv", "")] = summarise(p)
    print(json.dumps(res, indent=1))
