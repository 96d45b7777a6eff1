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


def build(force: bool = False, verbose: bool = False, trace: bool = False, defines=(), tag: str = "") -> str:
    """trace=True: developer build libgode_trace.so with -DGODE_TRACE (in-kernel time stamps; scripts/headline_trace.py loads it
    through GODE_LIB).  defines + tag: any other developer variant, libgode_<tag>.so with -D<define>... (e.g. --define
    GODE_ADJ_TIMING --tag adjtiming: per-phase clock64 accounting of the continuous adjoint kernel).  Never used by the product
    path."""
    if trace:
        defines, tag = tuple(defines) + ("GODE_TRACE",), tag or "trace"
    lib = LIB.replace("libgode.so", "libgode_{}.so".format(tag)) if tag else LIB
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= _deps_mtime():
        return lib
    nvcc = _nvcc()
    objdir = os.path.join(CSRC, "build_" + tag if tag else "build")
    os.makedirs(objdir, exist_ok=True)
    extra = ["-D" + d for d in defines]

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


TORCH_EXT = os.path.join(HERE, "_gode_torch.so")
TORCH_SRC = os.path.join(HERE, "csrc_torch", "gode_torch.cpp")


def build_torch_ext(force: bool = False) -> str:
    """Compile the thin PyTorch C++ host (csrc_torch/gode_torch.cpp) in-tree as gan_ode_b200/_gode_torch.so with g++ against
    this interpreter's torch (same image on the GPU box, so the prebuilt module travels).  It holds no kernels and links
    nothing from libgode.so (the entry points are bound by address at import), so it needs no nvcc."""
    deps = [TORCH_SRC, os.path.join(ROOT, "include", "gode.h"), __file__]
    if not force and os.path.exists(TORCH_EXT) and os.path.getmtime(TORCH_EXT) >= max(os.path.getmtime(d) for d in deps):
        return TORCH_EXT
    import sysconfig

    import torch
    from torch.utils import cpp_extension as ce
    cuda_home = os.environ.get("CUDA_HOME") or "/usr/local/cuda"
    inc = ce.include_paths() + [sysconfig.get_paths()["include"], os.path.join(cuda_home, "include"),
                                os.path.join(ROOT, "include")]
    libdir = ce.library_paths()[0]
    # The SYSTEM g++ (dynamic libstdc++, the one torch's libraries were linked against).  $CXX is deliberately not honoured:
    # in this image it points at /opt/gcc, which links libstdc++ statically into the module — two C++ runtimes in one
    # process, and the first exception thrown across the boundary (any TORCH_CHECK) segfaults.
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=_gode_torch",
           "-DTORCH_API_INCLUDE_EXTENSION_H", "-D_GLIBCXX_USE_CXX11_ABI={}".format(int(torch._C._GLIBCXX_USE_CXX11_ABI)),
           *["-I" + i for i in inc], TORCH_SRC, "-o", TORCH_EXT, "-L" + libdir, "-Wl,-rpath," + libdir,
           "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building _gode_torch.so failed:\n{}\n{}".format(r.stdout[-4000:], r.stderr[-4000:]))
    return TORCH_EXT


if __name__ == "__main__":
    argv = sys.argv[1:]
    defs = [argv[i + 1] for i, a in enumerate(argv) if a == "--define" and i + 1 < len(argv)]
    tags = [argv[i + 1] for i, a in enumerate(argv) if a == "--tag" and i + 1 < len(argv)]
    if defs and not tags:
        raise SystemExit("--define needs --tag NAME (the variant is written to csrc/libgode_NAME.so)")
    print(build(force="--force" in argv, verbose="-v" in argv, trace="--trace" in argv, defines=defs, tag=tags[0] if tags else ""))
    if "--trace" not in argv and not tags:
        print(build_torch_ext(force="--force" in argv))
