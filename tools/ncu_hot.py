#!/usr/bin/env python
"""List the SASS instructions of an ncu source page (ncu -i X.ncu-rep --page source --csv) with the most stall samples.

    ncu -i gpurun_out/x.ncu-rep --page source --csv > /tmp/src.csv; python tools/ncu_hot.py /tmp/src.csv [stall column] [top n]
"""
import csv
import sys


def main():
    path = sys.argv[1]
    col = sys.argv[2] if len(sys.argv) > 2 else "# Samples"
    top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
    rows = list(csv.reader(open(path)))
    hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
    i_src, i_s, i_c, i_ex = hdr.index("Source"), hdr.index("# Samples"), hdr.index(col), hdr.index("Instructions Executed")
    data = [r for r in rows if len(r) > max(i_c, i_s) and r[i_s].isdigit()]
    total = sum(int(r[i_s]) for r in data)
    print("instructions", len(data), "samples", total, "sum(%s)" % col, sum(int(r[i_c] or 0) for r in data))
    top = sorted(range(len(data)), key=lambda k: -int(data[k][i_c] or 0))[:top_n]
    for k in sorted(top):
        r = data[k]
        print("%5d  %-80s samples %6s  %s %6s  exec %s" % (k, r[i_src].strip()[:80], r[i_s], col, r[i_c], r[i_ex]))


if __name__ == "__main__":
    main()
