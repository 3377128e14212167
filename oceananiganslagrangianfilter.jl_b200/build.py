"""Builds liblfb200.so (hand-written CUDA for sm_100a + the C ABI) in-tree with nvcc.

nvcc cross-compiles without a GPU, so this runs in the CPU build container; the resulting .so is
git-ignored but travels to the GPU box with the repository snapshot.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "liblfb200.so")
SOURCES = ["lfb_api.cu", "lfb_generic.cu", "lfb_ztma.cu", "lfb_remap.cu"]
HEADERS = ["lfb_internal.h", "lfb_math.cuh", "lfb_zm_common.cuh", os.path.join(ROOT, "include", "lfb200.h")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-ffp-contract=off", "-ccbin", "/usr/bin/g++"]
CFLAGS += os.environ.get("LFB_NVCC_FLAGS", "").split()      # experiments, e.g. LFB_NVCC_FLAGS=-DZM_TY=7


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    headers = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + headers):
            cmd = [NVCC, *ARCH, *CFLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", s, "-o", o]
            if verbose:
                cmd += ["-Xptxas", "-v"]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            for log in ex.map(run, jobs):
                if verbose and log:
                    print(log, file=sys.stderr)
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-ccbin", "/usr/bin/g++", "-o", LIB, *objs, "-lnccl"]
        run(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
