"""Build recipe for libpgslam_b200.so (nvcc, sm_100a only, in-tree).

`python -m pgslam_b200.build` or `__graft_entry__.build()`.  The flags are part
of the numeric contract: -fmad=false keeps fp32 distances / transforms and the
fp64 solvers free of fused multiply-adds.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "lib", "libpgslam_b200.so")
SOURCES = ["cloud.cu", "sort.cu", "index.cu", "kdorder.cu", "knn.cu", "dense.cu", "filters.cu", "icp.cu", "api.cu", "modules.cpp", "cloud_io.cpp"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off,-fvisibility=hidden",
    "-ccbin", "/usr/bin/g++",
]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.sep not in c or os.path.exists(c)):
            return c
    raise RuntimeError("nvcc not found")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, "..", "include", "pgslam_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    os.makedirs(os.path.join(HERE, "lib"), exist_ok=True)
    objdir = os.path.join(HERE, "lib", "obj")
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    extra = os.environ.get("PGS_NVCC_EXTRA", "").split()  # tuning experiments only
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(objdir, src.rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {src} ---\n{out}\n")
        failed = failed or p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-cudart", "static", "-ccbin", "/usr/bin/g++",
                    "-Xlinker", "--exclude-libs,ALL"], check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
