"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` log:
one row per kernel launch of the LAST pass (the earlier passes are warm-up)."""
import csv
import sys


def main(path, passes=2):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    d = {}
    for r in rows[1:]:
        d.setdefault((int(r[ii]), r[ki].split("(")[0][:70]), {})[r[mi]] = float(r[vi].replace(",", ""))
    items = sorted(d.items())
    items = items[len(items) - len(items) // passes:]
    tot = 0.0
    print(f"| kernel | device µs | DRAM read MB | DRAM write MB | read GB/s |")
    print("|---|---|---|---|---|")
    for (_, k), m in items:
        us = m.get("gpu__time_duration.sum", 0) / 1e3
        rd, wr = m.get("dram__bytes_read.sum", 0) / 1e6, m.get("dram__bytes_write.sum", 0) / 1e6
        tot += us
        print(f"| `{k}` | {us:.1f} | {rd:.1f} | {wr:.1f} | {rd / us * 1e3 if us else 0:.0f} |")
    print(f"| **sum** | **{tot:.1f}** | | | |")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2)
