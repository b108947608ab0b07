"""Build every native artefact IN-TREE (they travel to the GPU box with gpurun, they are git-ignored).

    libmspack_b200/libmsgpu.so        CUDA kernels + C-ABI (nvcc, sm_100a)              - the product
    libmspack_b200/libmspack_dropin.so lzxd_* / qtmd_* / mszipd_* entry points on top   - the product
    libmspack_b200/gen/libmsgen.so    synthetic workload generator (gcc)
    oracle/_ref/*.so                  CPU oracles (gcc; the reference one only where /root/reference exists)
    tests/emul/libmsgpu_emul.so       host emulation of the device code for the CPU-only tests
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(s) and os.path.getmtime(s) > t for s in sources)


def _nvcc() -> str:
    for c in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def build_msgpu(force: bool = False) -> str:
    so = os.path.join(PKG, "libmsgpu.so")
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "msgpu.h"), os.path.join(ROOT, "include", "msgpu_cab.h"), os.path.join(ROOT, "include", "msgpu_chm.h")]
    if force or _newer(so, srcs):
        subprocess.check_call([_nvcc()] + NVCC_FLAGS + ["-o", so, os.path.join(CSRC, "msgpu.cu"), os.path.join(CSRC, "msgpu_cab.cu"), os.path.join(CSRC, "msgpu_chm.cu"),
                                                               os.path.join(CSRC, "msgpu_digest.cu")])
    return so


def build_dropin(force: bool = False) -> str | None:
    src = os.path.join(CSRC, "mspack_dropin.c")
    if not os.path.exists(src):
        return None
    so = os.path.join(PKG, "libmspack_dropin.so")
    if force or _newer(so, [src, os.path.join(ROOT, "include", "mspack_dropin.h"), os.path.join(ROOT, "include", "msgpu.h")]):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-I" + os.path.join(ROOT, "include"), "-o", so, src,
                               "-L" + PKG, "-lmsgpu", "-Wl,-rpath,$ORIGIN"])
    return so


def build_gen(force: bool = False) -> str:
    src = os.path.join(PKG, "gen", "msgen.c")
    so = os.path.join(PKG, "gen", "libmsgen.so")
    if force or _newer(so, [src]):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src, "-lpthread"])
    return so


def build_oracle() -> None:
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"])


def build_emul(force: bool = False) -> str:
    src = os.path.join(ROOT, "tests", "emul", "emul.cpp")
    so = os.path.join(ROOT, "tests", "emul", "libmsgpu_emul.so")
    srcs = [src] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    if force or _newer(so, srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src])
    return so


def build_all(force: bool = False) -> None:
    build_msgpu(force)
    build_dropin(force)
    build_gen(force)
    build_oracle()
    build_emul(force)


if __name__ == "__main__":
    build_all("--force" in sys.argv)
    print("built")
