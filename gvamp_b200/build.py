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
CU_SOURCES = ["capi.cu", "cg.cu", "layout.cu", "stats.cu", "matvec_tile.cu", "misslist.cu", "twin.cu", "assoc.cu", "people.cu", "vecops.cu"]
# test-only build (gvamp_b200/lib/xcheck/libgvamp_b200.so): the same library plus the two earlier kernel generations that the parity tests
# use as on-device cross-checks (env GVB_KERNELS=simple|lut1); capi.cu is compiled a second time with -DGVB_LEGACY_KERNELS for it
XCHECK_EXTRA = ["matvec_simple.cu", "matvec_lut.cu"]
XCHECK_LIB = os.path.join(LIBDIR, "xcheck", "libgvamp_b200.so")
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
    os.makedirs(os.path.dirname(XCHECK_LIB), exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [os.path.join(ROOT, "include", "gvamp_b200.h")]
    procs = []

    def compile_obj(src, obj, extra):
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, obj)
        if force or _newer(o, [s] + headers):
            cmd = [NVCC] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if ptxas_v else []) + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd)))
        return o

    objs = [compile_obj(src, src.replace(".cu", ".o"), []) for src in CU_SOURCES]
    xobjs = [o for o in objs if not o.endswith(os.sep + "capi.o")]
    xobjs.append(compile_obj("capi.cu", "capi_xcheck.o", ["-DGVB_LEGACY_KERNELS"]))
    xobjs += [compile_obj(src, src.replace(".cu", ".o"), ["-DGVB_LEGACY_KERNELS"]) for src in XCHECK_EXTRA]
    for src, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    # -Bsymbolic: a process may hold both builds (the tests do); each binds its own gvb_* symbols
    link = ARCH + ["-ccbin", HOSTCXX, "-lnccl", "-L/usr/lib/x86_64-linux-gnu", "-Xlinker", "-Bsymbolic"]
    if force or procs or _newer(LIB, objs):
        _run([NVCC, "-shared", "-o", LIB] + objs + link, verbose)
    if force or procs or _newer(XCHECK_LIB, xobjs):
        _run([NVCC, "-shared", "-o", XCHECK_LIB] + xobjs + link, verbose)
    return LIB


def build_host(force=False, verbose=True):
    if not os.path.isdir(HOST):
        return []
    os.makedirs(BINDIR, exist_ok=True)
    common = [os.path.join(HOST, f) for f in ("options.cpp", "data.cpp", "vamp.cpp", "utilities.cpp", "comm.cpp") if os.path.exists(os.path.join(HOST, f))]
    headers = [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith((".hpp", ".inc"))] + [os.path.join(ROOT, "include", "gvamp_b200.h")]
    outs = []
    cuda_inc = os.path.join(os.path.dirname(os.path.dirname(NVCC)), "include")    # nvtx3 (header-only) for the phase ranges
    flags = ["-O2", "-std=c++17", "-fPIC", "-I" + os.path.join(ROOT, "include"), "-I" + HOST, "-isystem", cuda_inc, "-Wno-unused-result"]
    link = ["-L" + LIBDIR, "-lgvamp_b200", "-Wl,-rpath," + LIBDIR, "-Wl,-rpath,$ORIGIN/../lib", "-lpthread", "-ldl"]
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
