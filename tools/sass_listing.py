#!/usr/bin/env python
"""Trimmed SASS listing of one kernel of libazb.so for profiles/: mnemonic histogram + every tensor-core / tensor-memory /
bulk-copy / mbarrier / reduction instruction with its address (what proves the tcgen05 + TMEM + bulk-copy claims without
rebuilding).   python tools/sass_listing.py <all.sass from cuobjdump -sass> <mangled function substring> <out.txt>"""
import collections
import re
import sys

KEEP = re.compile(r"\b(UTC\w+|LDTM\w*|STTM\w*|UBLKCP\w*|UTMA\w+|SYNCS\w*|RED\w*|REDUX\w*|LDGSTS\w*|ATOMS?\w*|MUFU\.\w+|F2FP\.\w+|BAR\.\w+|ACQBULK|UCGABAR\w*)")


def main():
    src, key, out = sys.argv[1], sys.argv[2], sys.argv[3]
    lines, on = [], False
    for ln in open(src):
        if "Function :" in ln:
            on = key in ln
            name = ln.split("Function :")[1].strip() if on else None
            if on:
                fname = name
        elif on and re.search(r"/\*[0-9a-f]{4,}\*/", ln):
            m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                lines.append((m.group(1), m.group(2).strip()))
    hist = collections.Counter()
    for _, ins in lines:
        op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
        hist[op.split(".")[0]] += 1
    with open(out, "w") as f:
        f.write("# %s\n# %d SASS instructions (cuobjdump -sass libazb.so, sm_100a)\n#\n# mnemonic histogram (base opcode: count)\n" % (fname, len(lines)))
        for op, c in sorted(hist.items(), key=lambda kv: -kv[1]):
            f.write("#   %-12s %5d\n" % (op, c))
        f.write("#\n# tensor-core (UTC*MMA = tcgen05.mma, UTCBAR = tcgen05.commit), tensor-memory (LDTM / STTM = tcgen05.ld / st), bulk-copy\n"
                "# (UBLKCP = cp.async.bulk), mbarrier (SYNCS), reduction / atomic, SFU and pack instructions, by address:\n")
        for addr, ins in lines:
            if KEEP.search(ins):
                f.write("/*%s*/  %s\n" % (addr, ins))
    print(out, len(lines), "instructions;", sum(1 for _, i in lines if KEEP.search(i)), "listed")


if __name__ == "__main__":
    main()
