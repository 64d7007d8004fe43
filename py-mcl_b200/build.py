"""Builds libmcl.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libmcl.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh"))] + \
           [os.path.join(os.path.dirname(HERE), "include", "mcl.h")]


def needs_build():
    if not os.path.isfile(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    # MCL_NVCC_EXTRA: extra compile flags for experiments, e.g. "-DMCL_TC_DIAG" (the MMA-warp wait probe read by MCL_TC9_DIAG=1)
    extra = os.environ.get("MCL_NVCC_EXTRA", "").split()
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, os.path.join(CSRC, "mcl.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        # with -Xptxas -v the tail of stderr is resource usage: show the diagnostics, not the last 4000 characters
        lines = [l for l in r.stderr.splitlines() if not l.startswith("ptxas info") and "bytes stack frame" not in l]
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), "\n".join(lines)[-6000:]))
    if verbose:
        print(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force=True, verbose=True))
