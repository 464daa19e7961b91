"""Build libhydrotools_b200.so in-tree with nvcc for sm_100a.

    python hydro-tools_b200/build.py [--force] [--verbose]

The shared library lands next to the Python host code
(hydrotools/cuda/_lib/libhydrotools_b200.so), is git-ignored, and travels to
the GPU box with the repo snapshot.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "hydrotools", "cuda", "_lib")
OUT = os.path.join(OUT_DIR, "libhydrotools_b200.so")
OBJ_DIR = os.path.join(HERE, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--fmad=true",
]


def _newer(src, dst, extra=()):
    if not os.path.isfile(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(s) > t for s in (src,) + tuple(extra))


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OUT_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    headers = tuple(glob.glob(os.path.join(CSRC, "*.cuh")) +
                    glob.glob(os.path.join(HERE, "..", "include", "*.h")))
    objs, procs = [], []
    for src in sorted(glob.glob(os.path.join(CSRC, "*.cu"))):
        obj = os.path.join(OBJ_DIR, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _newer(src, obj, headers):
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
                ["-c", src, "-o", obj]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE,
                                                stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            failed = True
            sys.stderr.write(f"nvcc failed on {src}\n")
    if failed:
        raise RuntimeError("nvcc compilation failed")
    if force or procs or not os.path.isfile(OUT):
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
               "-o", OUT] + objs + ["-cudart", "static"]
        subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
