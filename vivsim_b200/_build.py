"""Build recipe for libvivsim_b200.so (sm_100a only, in-tree).

    python -m vivsim_b200._build [--force] [--verbose]

nvcc cross-compiles without a GPU.  Objects land in vivsim_b200/csrc/build/, the shared
library next to this file so that it travels with the repository snapshot.
"""

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvivsim_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC",
         "-DVSB_BUILDING"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))   # csrc/xla/*.cc needs jaxlib headers: not built here


def _deps_mtime():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "vivsim_b200.h"))
    hdrs.append(os.path.abspath(__file__))
    return max(os.path.getmtime(h) for h in hdrs)


def build(force=False, verbose=False, out=OUT, defines=()):
    """`out` / `defines` build a tuning variant next to the product library (scripts/tune_variants.py)."""
    bdir = os.path.join(CSRC, "build") if out == OUT else out + ".obj"
    os.makedirs(bdir, exist_ok=True)
    dep = _deps_mtime()
    jobs = []
    objs = []
    for src in sources():
        obj = os.path.join(bdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        src_time = os.path.getmtime(src)
        if os.path.basename(src).startswith("vsb_step_part"):      # these include vsb_step.cu
            src_time = max(src_time, os.path.getmtime(os.path.join(CSRC, "vsb_step.cu")))
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(src_time, dep):
            cmd = [NVCC, *ARCH, *FLAGS, *["-D" + d for d in defines], "-c", src, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    failed = False
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            failed |= r.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed (see output above)")
    if jobs or force or not os.path.exists(out):
        cmd = [NVCC, *ARCH, "-shared", "-o", out, *objs, "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
