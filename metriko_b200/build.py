"""In-tree builds of the native libraries (no JIT cache: the .so files travel with the repo snapshot).

  libmetriko_b200.so     CUDA kernels + solver drivers + the C-ABI (include/metriko_b200.h), sm_100a only
  tests/hostglue/libmetriko_hostglue.so  test-input generator: host-side restatement of the reference's unchanged bookkeeping stages
                         (CPU only; NOT part of the product package -- it lives under tests/ and is linked by nothing in metriko_b200/)
"""
from __future__ import annotations

import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOSTGLUE = os.path.join(ROOT, "tests", "hostglue")
LIB_CUDA = os.path.join(PKG, "libmetriko_b200.so")
LIB_HOST = os.path.join(HOSTGLUE, "libmetriko_hostglue.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CUDA_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fopenmp", "-Xcompiler", "-O3", "-Xcompiler", "-fcx-limited-range",
]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _sources(d: str, exts: tuple[str, ...]) -> list[str]:
    out = []
    for base, _, files in os.walk(d):
        for f in sorted(files):
            if f.endswith(exts):
                out.append(os.path.join(base, f))
    return out


def build_hostglue(force: bool = False, verbose: bool = False) -> str:
    deps = _sources(HOSTGLUE, (".cpp", ".hpp"))
    if not force and _newer(LIB_HOST, deps):
        return LIB_HOST
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB_HOST,
           os.path.join(HOSTGLUE, "hostglue_capi.cpp")]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    subprocess.run(cmd, check=True)
    return LIB_HOST


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    """One nvcc -c per .cu in parallel (objects under metriko_b200/build/), then one link. `verbose` keeps the ptxas -v output
    (registers, spills, shared memory per kernel) in metriko_b200/build/<name>.ptxas.log."""
    from concurrent.futures import ThreadPoolExecutor

    deps = _sources(CSRC, (".cu", ".cuh", ".h", ".hpp", ".cpp")) + _sources(os.path.join(ROOT, "include"), (".h",))
    if not force and _newer(LIB_CUDA, deps):
        return LIB_CUDA
    srcs = _sources(CSRC, (".cu",))
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [d for d in deps if not d.endswith(".cu")]

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and _newer(obj, [src] + headers):
            return obj
        cmd = [NVCC, *CUDA_FLAGS, "-c", "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", obj, src]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose:
            with open(obj[:-2] + ".ptxas.log", "w") as fh:
                fh.write(r.stderr)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise subprocess.CalledProcessError(r.returncode, cmd)
        return obj

    with ThreadPoolExecutor(max_workers=max(1, min(len(srcs), os.cpu_count() or 1))) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_CUDA, *objs, "-lcudart", "-lgomp"]
    subprocess.run(cmd, check=True)
    return LIB_CUDA


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_hostglue(force, verbose)
    build_cuda(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
