#!/usr/bin/env python
"""Per-source-function instruction / stall breakdown of one kernel from an ncu --set full capture.

    python tools/sass_breakdown.py <report.ncu-rep> <lib.so> <mangled-kernel-substring> <units>

Joins the SASS page of the report (per-instruction executed counts and stall samples) with nvdisasm's line
info of the same kernel in the built library, and groups by the innermost source function
(needs -lineinfo at compile time and an unchanged build).  `units` divides the instruction counts
(e.g. warp-steps per launch) to print instructions per unit."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(REPO, "azul_deep_reinforcement_learning_b200", "csrc")


def main():
    rep, lib, kname, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    dis = None
    for f in sorted(os.listdir(tmp)):
        out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kname in out:
            dis = out.splitlines()
            break
    assert dis, "kernel not found"
    start = [i for i, l in enumerate(dis) if l.startswith("//---") and ".text." in l and kname in l][0]
    lines, curfile, cur = [], None, None
    for l in dis[start + 1:]:
        if l.startswith("//---") and ".text." in l:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            curfile, cur = os.path.basename(m.group(1)), int(m.group(2))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
            lines.append((curfile, cur))
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    # the page holds one section per profiled launch ("Kernel Name" row, header row, one row per instruction):
    # take the first section of the wanted kernel
    want = re.sub(r"I?L?i(\d+)E?", "", kname).split("I")[0].lstrip("_Z0123456789").rstrip("E")
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    sec = next(((a, b) for a, b in zip(starts, starts[1:]) if (want in rows[a][1] or want.split("8")[-1] in rows[a][1]) and b - a - 2 == len(lines)), None)
    assert sec, ("no section of %d instructions for %s" % (len(lines), want), [(rows[a][1][:40], b - a - 2) for a, b in zip(starts, starts[1:])])
    hdr = rows[sec[0] + 1]
    idx = {h: i for i, h in enumerate(hdr)}
    data = rows[sec[0] + 2:sec[1]]
    assert len(data) == len(lines), (len(data), len(lines), "rebuild the library at the profiled commit")
    src = {}
    for fn in (f for f in os.listdir(CSRC) if os.path.isfile(os.path.join(CSRC, f))):
        src[fn] = open(os.path.join(CSRC, fn)).read().splitlines()

    def func_of(f, ln):
        if f not in src or not ln:
            return f or "?"
        name = f
        for i, l in enumerate(src[f][:ln]):
            m = re.match(r"\s*(?:AZB_HD|AZB_M|__device__|__global__|template).*?\b(\w+)\(", l)
            if m and not l.strip().startswith("//"):
                name = m.group(1)
        return name

    by_line = os.environ.get("BY_LINE")          # BY_LINE=azb_policy.cu: group that file's instructions by source line
    I, S, T = defaultdict(float), defaultdict(float), defaultdict(float)
    for (f, ln), r in zip(lines, data):
        k = "%s:%d %s" % (f, ln, src[f][ln - 1].strip()[:60]) if by_line and f == by_line else func_of(f, ln)
        I[k] += float(r[idx["Instructions Executed"]] or 0)
        T[k] += float(r[idx["Thread Instructions Executed"]] or 0)
        S[k] += float(r[idx["Warp Stall Sampling (All Samples)"]] or 0)
    ti, ts = sum(I.values()), sum(S.values())
    print("%-28s %7s %7s %6s %10s" % ("function", "inst%", "stall%", "lanes", "inst/unit"))
    for k in sorted(I, key=I.get, reverse=True)[:60 if by_line else 24]:
        print("%-28s %6.1f%% %6.1f%% %6.1f %10.1f" % (k if by_line else k[:28], 100 * I[k] / ti, 100 * S[k] / ts, T[k] / max(I[k], 1), I[k] / units))
    print("total instructions per unit: %.1f" % (ti / units))


if __name__ == "__main__":
    main()
