"""Builds proxytransformation_b200/csrc/*.cu into libptpreshape.so (in-tree) with nvcc for sm_100a.

Plain ``nvcc -shared``: the library exposes only the C ABI of include/pt_preshape.h (no torch types), links the
CUDA runtime statically and never links libcuda directly (driver entry points are resolved at run time), so it also
loads on a machine without a GPU driver.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libptpreshape.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
              "-Xcompiler", "-fPIC", "-cudart", "static"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "pt_preshape.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    extra = os.environ.get("PT_NVCC_DEFINES", "").split()       # debug builds only, e.g. PT_NVCC_DEFINES=-DPT_POOL_EVENTS
    if not force and not extra and not needs_build():
        return LIB
    objs = []
    procs = []
    for src in sources():
        obj = src[:-3] + ".o"
        objs.append(obj)
        cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != "--use_fast_math=false"] + (["-Xptxas", "-v"] if verbose else []) + extra + ["-c", src, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [_nvcc(), "-shared", "-cudart", "static", "-o", LIB] + objs
    subprocess.run(link, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
