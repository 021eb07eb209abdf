"""Builds hymd_b200/lib/libhymd_b200.so (sm_100a only) with nvcc.

    python -m hymd_b200.build [--force] [--verbose]

In-tree build: the shared library travels with the repository snapshot to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libhymd_b200.so")
SOURCES = ["context.cu", "sort.cu", "paint.cu", "kspace.cu", "readout.cu", "energy.cu",
           "slabfft.cu", "comm.cu", "migrate.cu", "xline.cu", "planefft.cu", "bonded.cu", "md.cu", "gpe.cu", "graph.cu"]
HEADERS = [os.path.join(CSRC, "ctx.cuh"), os.path.join(CSRC, "fft.cuh"), os.path.join(CSRC, "bonded.cuh"),
           os.path.join(CSRC, "md.cuh"), os.path.join(CSRC, "gpe.cuh"), os.path.join(HERE, "..", "include", "hymd_b200.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
              "-Xptxas", "-v"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libhymd_b200.so cannot be built")
    return nvcc


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + HEADERS):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        cmd = [nvcc] + ARCH + NVCC_FLAGS + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(o + ".log", "w") as fh:
            fh.write(" ".join(cmd) + "\n" + log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}:\n{log}")
        if verbose:
            print(log)
        return o

    with ThreadPoolExecutor(max_workers=min(6, os.cpu_count() or 1)) as ex:
        list(ex.map(compile_one, jobs))
    objs = [os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cuda_lib = os.path.join(os.path.dirname(os.path.dirname(nvcc)), "lib64")
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + [
            "-L" + cuda_lib, "-lcufft", "-ldl", "-Xlinker", "-rpath," + cuda_lib]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
