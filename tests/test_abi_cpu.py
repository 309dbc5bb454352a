"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/gvamp_b200.h declares, fails loudly (no CPU fallback) without a device, and the host-side
partition rule matches the reference's divide_work."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def capi():
    from gvamp_b200 import build, capi as C
    build.build_cuda(verbose=False)
    return C


def test_header_symbols_exported(capi):
    hdr = open(os.path.join(ROOT, "include", "gvamp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gvb_[a-zA-Z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 40
    lib = ctypes.CDLL(capi.LIB_PATH)
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"
    # and the Python binding covers the whole header
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)


def test_header_cites_reference(capi):
    hdr = open(os.path.join(ROOT, "include", "gvamp_b200.h")).read()
    for anchor in ("data.cpp:848-1011", "data.cpp:728-835", "data.cpp:392-485", "vamp.cpp:1130-1229", "vamp.cpp:805-869",
                   "utilities.cpp:259-291", "data.cpp:201-234", "vamp_probit.cpp:661-726"):
        assert anchor in hdr


def test_divide_work_matches_reference_rule(capi, oracle):
    for Mt, n in ((20000, 2), (2200000, 8), (1003, 8), (7, 8), (500000, 3)):
        tot = 0
        for r in range(n):
            M, S = capi.divide_work(Mt, n, r)
            assert (M, S) == oracle.divide_work(Mt, n, r)
            assert S == tot
            tot += M
        assert tot == Mt


def test_no_cpu_fallback(capi):
    """Without a GPU the context cannot be created: the product path must fail, never emulate."""
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.GvbError) as e:
        capi.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_library_ships_one_kernel_generation(capi):
    """The product library holds only the tile kernels; the earlier generations (FP64 decode kernels, first table kernels) that the
    parity tests use as on-device cross-checks exist only in the test build gvamp_b200/lib/xcheck/libgvamp_b200.so, which exports
    the same C ABI."""
    import subprocess
    syms = lambda path: subprocess.run(["nm", "-C", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    product, xcheck = syms(capi.LIB_PATH), syms(capi.XCHECK_LIB_PATH)
    for legacy in ("gvb_ax_lut", "gvb_atx_lut", "gvb_ax_simple", "gvb_atx_simple"):
        assert legacy not in product and legacy in xcheck, legacy
    assert "gvb_ax_tile" in product and "gvb_atx_tile" in product
    hdr = open(os.path.join(ROOT, "include", "gvamp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    lib = ctypes.CDLL(capi.XCHECK_LIB_PATH, mode=ctypes.RTLD_LOCAL)
    assert all(hasattr(lib, n) for n in set(re.findall(r"\b(gvb_[a-zA-Z0-9_]+)\s*\(", hdr)))
