#!/usr/bin/env python
"""Per source line totals (instructions executed, stall samples) of an `ncu --page source --csv --print-source cuda,sass` dump."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    cur_file = ""
    agg = {}
    hdr = None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if len(r) >= 2 and r[0] == "Function Name":
            continue
        if r and r[0] == "Line No":
            hdr = r
            ie = hdr.index("Instructions Executed"); sm = hdr.index("# Samples")
            continue
        if hdr is None or len(r) <= ie:
            continue
        if r[0] != "":  # a source line row: carries the line's totals
            try:
                key = (cur_file, int(r[0]))
                agg[key] = (int(r[ie]), int(r[sm]), r[1].strip())
            except ValueError:
                pass
    tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
    print(f"total instructions {tot_i}  samples {tot_s}")
    for (f, ln), (i, s, src) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{i:11d} {100 * i / max(1, tot_i):5.1f}%  smp {100 * s / max(1, tot_s):5.1f}%  {f}:{ln}  {src[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
