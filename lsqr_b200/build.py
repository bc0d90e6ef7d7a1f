"""In-tree build of the CUDA engine: nvcc -> lsqr_b200/lib/liblsqr_b200.so (sm_100a only).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# LSQR_B200_BUILD_TAG / LSQR_B200_EXTRA_NVCC_FLAGS build an experimental copy next to the product library
# (lib/liblsqr_b200.<tag>.so, e.g. with -DLSQRB_WARP_MINBLOCKS=3); run with LSQR_B200_LIB pointing at it.
TAG = os.environ.get("LSQR_B200_BUILD_TAG", "")
EXTRA = os.environ.get("LSQR_B200_EXTRA_NVCC_FLAGS", "").split()
OBJ = os.path.join(HERE, "build" + ("_" + TAG if TAG else ""))
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "liblsqr_b200" + ("." + TAG if TAG else "") + ".so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
         "-Xptxas", "-v", "--expt-relaxed-constexpr"]
SOURCES = ["engine.cu", "build_csr.cu", "synth.cu"]


def _deps(src: str) -> list[str]:
    deps = [os.path.join(CSRC, src)]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "lsqr_b200.h"))
    return deps


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    objs = []
    for src in SOURCES:
        if not os.path.exists(os.path.join(CSRC, src)):
            continue
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, _deps(src)):
            cmd = [NVCC, *ARCH, *FLAGS, *EXTRA, "-I", os.path.join(os.path.dirname(HERE), "include"),
                   "-c", os.path.join(CSRC, src), "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            log = r.stdout + r.stderr
            with open(obj + ".log", "w") as f:
                f.write(" ".join(cmd) + "\n" + log)
            if r.returncode != 0:
                sys.stderr.write(log)
                raise RuntimeError(f"nvcc failed on {src}")
            if verbose:
                print(log)
    if force or _stale(LIB, objs):
        cmd = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
