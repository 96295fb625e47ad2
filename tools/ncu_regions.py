#!/usr/bin/env python
"""Attribute the SASS-level samples / executed instructions of an ncu source page to CUDA source lines via nvdisasm -g.

    ncu -i X.ncu-rep --page source --csv > src.csv ; nvdisasm -g X.cubin > all.dis
    python tools/ncu_regions.py src.csv all.dis <mangled function> [file.cu:lo-hi=name ...]
Without region arguments prints the top source lines."""
import collections
import csv
import re
import sys


def main():
    src, dis, fun = sys.argv[1:4]
    regions = []
    for a in sys.argv[4:]:
        m = re.match(r"([^:]+):(\d+)-(\d+)=(.*)", a)
        regions.append((m.group(1), int(m.group(2)), int(m.group(3)), m.group(4)))
    cur, on, addr2line = None, False, {}
    for ln in open(dis):
        if ln.startswith(".text."):
            on = ln.strip().rstrip(":") == ".text." + fun
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
        if m and cur:
            addr2line[int(m.group(1), 16)] = cur
    rows = list(csv.reader(open(src)))
    hdr = next(r for r in rows if "Source" in r and "# Samples" in r)
    ia, i_s, i_ex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    data = [r for r in rows if len(r) > i_s and r[i_s].isdigit()]
    base = int(data[0][ia], 16)
    agg = collections.defaultdict(lambda: [0, 0])
    for r in data:
        k = addr2line.get(int(r[ia], 16) - base, ("?", 0))
        agg[k][0] += int(r[i_ex]); agg[k][1] += int(r[i_s])
    tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
    if regions:
        bag = collections.OrderedDict((r[3], [0, 0]) for r in regions)
        bag["(other)"] = [0, 0]
        for (f, l), v in agg.items():
            name = next((r[3] for r in regions if r[0] == f and r[1] <= l <= r[2]), "(other) " + f)
            b = bag.setdefault(name, [0, 0]); b[0] += v[0]; b[1] += v[1]
        for k, v in bag.items():
            if v[0]:
                print("%-44s instr %5.1f%%  samples %5.1f%%" % (k, 100 * v[0] / tot, 100 * v[1] / ts))
    else:
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
            print("%-22s:%-5d instr %5.1f%%  samples %5.1f%%" % (k[0], k[1], 100 * v[0] / tot, 100 * v[1] / ts))
    print("total warp instructions %d, samples %d" % (tot, ts))


if __name__ == "__main__":
    main()
