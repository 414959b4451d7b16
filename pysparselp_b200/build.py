"""In-tree build of ``csrc/libcpppd.so`` with nvcc for sm_100a (B200).

``python -m pysparselp_b200.build`` or ``build_library()``.  nvcc cross-compiles without
a GPU, so this also works on the CPU-only build box.  The library links the CUDA runtime
statically and ``dlopen``s NCCL lazily, so it has no load-time dependency besides libc.
"""
import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "csrc", "cpppd.cu")
OUT = os.path.join(_HERE, "csrc", "libcpppd.so")

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false",  # the reference's numpy/scipy arithmetic never fuses a*b+c
    "-shared", "-Xcompiler", "-fPIC",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def is_stale():
    if not os.path.isfile(OUT):
        return True
    newest = max(os.path.getmtime(os.path.join(root, f))
                 for root in (os.path.join(_HERE, "csrc"), os.path.join(_HERE, "..", "include"))
                 for f in os.listdir(root) if f.endswith((".cu", ".cuh", ".h")))
    return newest > os.path.getmtime(OUT)


def build_library(force=False, verbose=False):
    if not force and not is_stale():
        return OUT
    defines = os.environ.get("CPPPD_NVCC_DEFINES", "").split()  # e.g. "-DCPPPD_GATHER_CHUNK=3 -DCPPPD_MIN_BLOCKS=6"
    cmd = [find_nvcc()] + NVCC_FLAGS + defines + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC, "-ldl"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr))
    if verbose:
        print(res.stderr)
    return OUT


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
