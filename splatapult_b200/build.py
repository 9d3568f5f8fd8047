"""In-tree build of libsplat_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m splatapult_b200.build [--force] [--verbose]

Every .cu is compiled with -gencode arch=compute_100a,code=sm_100a -lineinfo.  project.cu additionally gets
-fmad=false (its records are bit-comparable with the CPU oracle); host code is built with -ffp-contract=off.
The .so stays inside the package directory so it travels with the repo snapshot (it is git-ignored).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libsplat_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
HOST_FLAGS = "-fPIC,-ffp-contract=off,-fvisibility=hidden,-Wall,-Wextra,-Wno-unused-parameter"
NVCC_COMMON = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", HOST_FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC,
               "--expt-relaxed-constexpr", "-Xptxas", "-v"]
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-fvisibility=hidden", "-Wall", "-Wextra",
             "-I", os.path.join(ROOT, "include"), "-I", CSRC]

CU_SOURCES = {
    "pack.cu": [],
    "project.cu": ["-fmad=false"],
    "plydecode.cu": ["-fmad=false"],
    "radix_sort.cu": [],
    "binsort.cu": [],
    "composite.cu": [],
    "renderer.cu": [],
    "multigpu.cu": [],
    "capi.cu": [],
}
CPP_SOURCES = ["ply.cpp", "gaussiancloud.cpp", "camerasconfig.cpp"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _deps_hash(extra: list[str]) -> str:
    h = hashlib.sha256()
    for d in (CSRC, os.path.join(ROOT, "include"), os.path.join(ROOT, "include", "splatapult_b200")):
        for name in sorted(os.listdir(d)):
            p = os.path.join(d, name)
            if os.path.isfile(p):
                h.update(name.encode())
                h.update(open(p, "rb").read())
    h.update(open(os.path.abspath(__file__), "rb").read())
    h.update(" ".join(extra).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    stamp = os.path.join(OBJ, "stamp")
    # diagnostic build: SPL_BUILD_DEFINES="SPL_SORT_TIMELINE" adds -D flags (see radix_sort.cu); never set by default
    defines = ["-D" + d for d in os.environ.get("SPL_BUILD_DEFINES", "").split() if d]
    digest = _deps_hash(defines)
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    nvcc = _nvcc()
    objs = []
    log = []
    for src, extra in CU_SOURCES.items():
        o = os.path.join(OBJ, src + ".o")
        cmd = [nvcc] + ARCH + NVCC_COMMON + extra + defines + ["-c", os.path.join(CSRC, src), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
        if r.returncode:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        objs.append(o)
    for src in CPP_SOURCES:
        if not os.path.exists(os.path.join(CSRC, src)):
            continue
        o = os.path.join(OBJ, src + ".o")
        cmd = ["g++"] + CXX_FLAGS + ["-c", os.path.join(CSRC, src), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
        if r.returncode:
            raise RuntimeError(f"g++ failed for {src}:\n{r.stdout}\n{r.stderr}")
        objs.append(o)
    cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
    if r.returncode:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(os.path.join(OBJ, "build.log"), "w") as f:
        f.write("\n".join(log))
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
