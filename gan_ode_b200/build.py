"""Build libgode.so in-tree with nvcc for sm_100a (B200).  No torch headers: the library is plain C ABI.

    python -m gan_ode_b200.build            # incremental
    python -m gan_ode_b200.build --force
"""
from __future__ import annotations

import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(CSRC, "libgode.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; set NVCC=/path/to/nvcc")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime() -> float:
    newest = 0.0
    for d in (CSRC, os.path.join(ROOT, "include")):
        for f in os.listdir(d):
            if f.endswith((".cu", ".cuh", ".h")):
                newest = max(newest, os.path.getmtime(os.path.join(d, f)))
    return max(newest, os.path.getmtime(__file__))


def build(force: bool = False, verbose: bool = False, trace: bool = False) -> str:
    """trace=True: developer build libgode_trace.so with -DGODE_TRACE (in-kernel time stamps; scripts/headline_trace.py loads it
    through GODE_LIB).  Never used by the product path."""
    lib = LIB.replace("libgode.so", "libgode_trace.so") if trace else LIB
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= _deps_mtime():
        return lib
    nvcc = _nvcc()
    objdir = os.path.join(CSRC, "build_trace" if trace else "build")
    os.makedirs(objdir, exist_ok=True)
    extra = ["-DGODE_TRACE"] if trace else []

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for {}:\n{}\n{}".format(src, r.stdout, r.stderr))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(compile_one, sources()))
    r = subprocess.run([nvcc, "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n{}\n{}".format(r.stdout, r.stderr))
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv))
