#!/usr/bin/env python
"""Variant builds of libazb.so for tuning sweeps:  python tools/build_variant.py NAME -DAZB_STEP_CLAIM=2 ...
Recompiles csrc/azb.cu (or the unit named by UNIT=azb_policy) with the extra flags and links it with the default build's
other objects into gpurun_variants/NAME.so (git-ignored; travels with the snapshot).  Load it with AZB_LIB=gpurun_variants/NAME.so."""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from azul_deep_reinforcement_learning_b200 import build as B  # noqa: E402


def main():
    name, extra = sys.argv[1], sys.argv[2:]
    unit = os.environ.get("UNIT", "azb")
    B.build()
    out_dir = os.path.join(REPO, "gpurun_variants")
    os.makedirs(out_dir, exist_ok=True)
    nvcc = B.find_nvcc()
    flags = [f for f in B.NVCC_FLAGS if f != "-shared"]
    obj = os.path.join(out_dir, name + "." + unit + ".o")
    subprocess.check_call([nvcc] + flags + extra + ["-c", "-o", obj, os.path.join(B.CSRC, unit + ".cu")])
    objs = [obj if os.path.basename(s)[:-3] == unit else os.path.join(B.CSRC, "_obj", os.path.basename(s)[:-3] + ".o") for s in B.SOURCES]
    so = os.path.join(out_dir, name + ".so")
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", so] + objs)
    os.remove(obj)
    print(so)


if __name__ == "__main__":
    main()
