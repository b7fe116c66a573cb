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


if __name__ == "__main__":
    print(build(force=True))
