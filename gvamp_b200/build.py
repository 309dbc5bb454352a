"""Build libgvamp_b200.so (CUDA, sm_100a only) and the C++ host programs in-tree.

    python -m gvamp_b200.build [--force]

nvcc cross-compiles without a GPU.  The shared library lands in gvamp_b200/lib/ (git-ignored, but it
travels to the GPU box with the snapshot).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIBDIR = os.path.join(HERE, "lib")
BINDIR = os.path.join(HERE, "bin")
LIB = os.path.join(LIBDIR, "libgvamp_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOSTCXX = "/usr/bin/g++"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CU_SOURCES = ["capi.cu", "cg.cu", "layout.cu", "stats.cu", "matvec_simple.cu", "matvec_lut.cu", "matvec_tile.cu", "misslist.cu", "twin.cu", "assoc.cu", "people.cu", "vecops.cu"]
NVCC_FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-ccbin", HOSTCXX, "--expt-relaxed-constexpr",
              "-Xcompiler", "-Wno-unused-result"] + ARCH


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)


def build_cuda(force=False, verbose=True, ptxas_v=False):
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [os.path.join(ROOT, "include", "gvamp_b200.h")]
    objs, procs = [], []
    for src in CU_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + headers):
            cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if ptxas_v else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd)))
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    if force or procs or _newer(LIB, objs):
        _run([NVCC, "-shared", "-o", LIB] + objs + ARCH + ["-ccbin", HOSTCXX, "-lnccl", "-L/usr/lib/x86_64-linux-gnu"], verbose)
    return LIB


def build_host(force=False, verbose=True):
    if not os.path.isdir(HOST):
        return []
    os.makedirs(BINDIR, exist_ok=True)
    common = [os.path.join(HOST, f) for f in ("options.cpp", "data.cpp", "vamp.cpp", "utilities.cpp", "comm.cpp") if os.path.exists(os.path.join(HOST, f))]
    headers = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".hpp")] + [os.path.join(ROOT, "include", "gvamp_b200.h")]
    outs = []
    flags = ["-O2", "-std=c++17", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + HOST, "-Wno-unused-result"]
    link = ["-L" + LIBDIR, "-lgvamp_b200", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath,$ORIGIN/../lib", "-lpthread"]
    hostlib = os.path.join(LIBDIR, "libgvamp_host.so")
    capi = os.path.join(HOST, "host_capi.cpp")
    if common and os.path.exists(capi) and (force or _newer(hostlib, common + [capi] + headers + [LIB])):
        _run([HOSTCXX] + flags + ["-shared", "-o", hostlib] + common + [capi] + link, verbose)
    if os.path.exists(hostlib):
        outs.append(hostlib)
    for main in ("main_real", "main_real_probit"):
        src = os.path.join(HOST, main + ".cpp")
        exe = os.path.join(BINDIR, main)
        if os.path.exists(src) and common:
            if force or _newer(exe, common + [src] + headers + [LIB]):
                _run([HOSTCXX] + flags + ["-o", exe, src] + common + link, verbose)
            outs.append(exe)
    return outs


def build_all(force=False, verbose=True):
    lib = build_cuda(force, verbose)
    return [lib] + build_host(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
