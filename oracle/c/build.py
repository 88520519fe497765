"""Compile the oracle's C restatement (gcc -O3 -fopenmp) into oracle/c/libiblbm_ref.so (fp32, what every parity test and
the CPU baseline use) and oracle/c/libiblbm_ref64.so (the same source with REF_REAL=double: the fp64 yardstick).
Test infrastructure."""

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "iblbm_ref.c")
OUT = os.path.join(HERE, "libiblbm_ref.so")
OUT64 = os.path.join(HERE, "libiblbm_ref64.so")


def _compile(out, defines, force):
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-fno-fast-math", "-ffp-contract=off", *defines,
                        "-shared", "-fPIC", SRC, "-o", out, "-lm"], check=True)
    return out


def build(force=False):
    _compile(OUT64, ["-DREF_REAL=double"], force)
    return _compile(OUT, [], force)


def build64(force=False):
    return _compile(OUT64, ["-DREF_REAL=double"], force)


if __name__ == "__main__":
    print(build(force=True))
