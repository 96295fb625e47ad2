"""In-tree build of the CUDA library (``libazb.so``) for sm_100a.

``nvcc`` cross-compiles without a GPU, so this runs in the build container; the resulting ``.so``
is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libazb.so")
SOURCES = [os.path.join(CSRC, "azb.cu"), os.path.join(CSRC, "azb_policy.cu"), os.path.join(CSRC, "azb_a2c.cu"), os.path.join(CSRC, "azb_update.cu"), os.path.join(CSRC, "azb_variant.cu")]
HEADERS = [os.path.join(CSRC, "azb_rules.cuh"), os.path.join(CSRC, "azb_tc.cuh"), os.path.join(CSRC, "azb_variant.cuh"), os.path.join(CSRC, "azb_internal.h"), os.path.join(CSRC, "azb_queue.cuh"), os.path.join(os.path.dirname(PKG_DIR), "include", "azb.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",   # B200 only: no other arch, no PTX fallback
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(p) > t for p in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile libazb.so if missing or stale; returns its path."""
    if force or needs_build():
        cmd = [find_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + SOURCES
        subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
