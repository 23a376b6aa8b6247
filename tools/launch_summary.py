#!/usr/bin/env python
"""Summarise an `ncu --csv` launch list: per kernel (name, grid) mean duration and DRAM bytes."""
import collections
import csv
import sys


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    d = collections.defaultdict(lambda: collections.defaultdict(list))
    for row in csv.DictReader(lines):
        d[row["Kernel Name"][:44] + " " + row["Grid Size"]][row["Metric Name"]].append(float(row["Metric Value"]))
    print(f"{'kernel (grid)':62s} {'n':>4s} {'us':>8s} {'rd MB':>8s} {'wr MB':>8s} {'GB/s':>8s} {'Minst':>7s}")
    for k, m in sorted(d.items(), key=lambda kv: -sum(kv[1]["gpu__time_duration.sum"])):
        t = m["gpu__time_duration.sum"]
        n = len(t)
        rd = sum(m.get("dram__bytes_read.sum", [0])) / n
        wr = sum(m.get("dram__bytes_write.sum", [0])) / n
        tm = sum(t) / n
        inst = sum(m.get("smsp__inst_executed.sum", [0])) / n / 1e6
        print(f"{k:62s} {n:4d} {tm / 1000:8.2f} {rd / 1e6:8.2f} {wr / 1e6:8.2f} {(rd + wr) / tm:8.1f} {inst:7.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
