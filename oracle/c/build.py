"""Compile the oracle's C restatement (gcc -O3 -fopenmp) into oracle/c/libiblbm_ref.so.  Test infrastructure."""

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "iblbm_ref.c")
OUT = os.path.join(HERE, "libiblbm_ref.so")


def build(force=False):
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fno-fast-math", "-shared", "-fPIC", SRC, "-o", OUT, "-lm"],
                       check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
