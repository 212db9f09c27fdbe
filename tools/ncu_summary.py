#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep into the small JSON bench.py's roofline block cites.

    python tools/ncu_summary.py gpurun_out/fused.ncu-rep wmd_fused_small_kernel 65536 profiles/r02_ncu_fused.json
"""
import csv
import json
import subprocess
import sys


def main(rep, kernel, pairs_per_launch, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    pick = [r for r in rows[2:] if kernel in r[hdr.index("Kernel Name")]]
    if not pick:
        raise SystemExit(f"no launch of {kernel} in {rep}")
    r = pick[len(pick) // 2]                                   # a launch from the middle of the run
    g = lambda name: float(r[hdr.index(name)].replace(",", "")) if name in hdr and r[hdr.index(name)] else None
    unit = lambda name: rows[1][hdr.index(name)] if name in hdr else ""
    to_bytes = lambda name: (g(name) or 0.0) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit(name), 1)
    inst = g("smsp__inst_executed.sum")
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    d = {"source": f"ncu --set full --clock-control none, kernel {kernel}, launch {len(pick) // 2} of {len(pick)}",
         "commit": commit, "pairs_per_launch": int(pairs_per_launch),
         "duration_us": (g("gpu__time_duration.sum") or 0.0) / (1e3 if unit("gpu__time_duration.sum") == "ns" else 1),
         "dram_bytes_per_launch": to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"),
         "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
         "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
         "warp_instructions": inst, "warp_instructions_per_pair": inst / float(pairs_per_launch) if inst else None,
         "registers_per_thread": g("launch__registers_per_thread"),
         "fp64_pipe_pct": g("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
         "alu_pipe_pct": g("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
         "l2_throughput_pct": g("lts__throughput.avg.pct_of_peak_sustained_elapsed")}
    json.dump(d, open(out, "w"), indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4])
