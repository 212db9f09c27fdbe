#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K > dump.csv
    python tools/ncu_lines.py dump.csv [top_n]
Prints, per source line: warp instructions executed, share, stall samples and the top stall reasons.
"""
import csv
import sys


SORT = 1


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur_file = ""
    hdr = None
    out = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] == "":
            continue
        d = dict(zip(hdr[4:], r[4:]))
        try:
            inst = int(d["Instructions Executed"])
            samp = int(d["# Samples"])
        except (KeyError, ValueError):
            continue
        stalls = {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit() and int(v)}
        out.append((inst, samp, cur_file, r[0], r[1].strip(), stalls))
    tot_i = sum(o[0] for o in out) or 1
    tot_s = sum(o[1] for o in out) or 1
    print(f"total warp instructions {tot_i}, samples {tot_s}")
    for inst, samp, f, ln, src, st in sorted(out, key=lambda o: -o[SORT])[:top]:
        top3 = ", ".join(f"{k[6:]}:{v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(f"{100*inst/tot_i:5.1f}%i {100*samp/tot_s:5.1f}%s  {f}:{ln:>4}  {src[:70]:70s} | {top3}")


if __name__ == "__main__":
    if len(sys.argv) > 3 and sys.argv[3] == "inst":
        SORT = 0
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
