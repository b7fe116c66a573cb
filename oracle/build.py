"""ORACLE — test infrastructure only (see oracle/__init__.py).

Compiles oracle/geom.c (the C restatement of pytorch3d's CPU ball query / FPS)
into oracle/_build/libptoracle_geom.so with gcc.  ``-ffp-contract=off`` and no
``-march=native`` so that ``((dx*dx)+(dy*dy))+(dz*dz)`` is evaluated without FMA
contraction, as upstream's CPU build and torch's elementwise kernels do
(SURVEY.md §8c pinning decision iv).
"""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libptoracle_geom.so")
SRC = os.path.join(HERE, "geom.c")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force) and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


# ---- oracle/_ref: the reference's own hot-path module, compiled where it lies ------------------------------------------
# The reference is pure Python (setup.py:108 ext_modules=[]): "compiling" it means byte-compiling the UNMODIFIED file
# /root/reference/embodiedscan/models/necks/preshape_norm_reverse_drop.py into oracle/_ref/ (git-ignored, not gpurun-ignored,
# so the bytecode travels to the GPU box like a built .so — under a neutral extension: snapshot tools drop *.pyc; no reference source
# enters the repository).  oracle/ref_shim.py
# loads it there under the same three import shims, which makes the CPU arm of bench.py the reference itself
# (cpu_baseline.kind = "reference") instead of the restatement.
REF_SRC = "/root/reference/embodiedscan/models/necks/preshape_norm_reverse_drop.py"
REF_DIR = os.path.join(HERE, "_ref")
REF_PYC = os.path.join(REF_DIR, "preshape_norm_reverse_drop.bytecode")


def build_ref(force: bool = False):
    """Byte-compile the reference module into oracle/_ref/ when /root/reference is present (build container); returns the path
    of the bytecode file, or None when neither the source nor a previously built file exists."""
    import py_compile
    if os.path.exists(REF_SRC):
        os.makedirs(REF_DIR, exist_ok=True)
        if force or not os.path.exists(REF_PYC) or os.path.getmtime(REF_PYC) < os.path.getmtime(REF_SRC):
            py_compile.compile(REF_SRC, cfile=REF_PYC, dfile="embodiedscan/models/necks/preshape_norm_reverse_drop.py", doraise=True)
    return REF_PYC if os.path.exists(REF_PYC) else None


if __name__ == "__main__":
    print(build(force=True))
    print(build_ref(force=True))
