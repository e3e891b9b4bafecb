"""CPU-only checks of the drop-in boundary: the library builds/loads and exports exactly the symbols that
include/h264b200.h declares; without a GPU the product refuses to work instead of falling back to a CPU path."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "h264b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(h264b_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_all_exported():
    from h264decode_b200 import build
    lib = build.build()
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib], text=True)
    exported = sorted({l.split()[-1] for l in out.splitlines() if " T " in l and "h264b_" in l})
    decl = declared_symbols()
    assert decl == exported
    from h264decode_b200 import capi
    assert sorted(capi.SYMBOLS) == decl
    L = ctypes.CDLL(lib)
    for s in decl:
        assert getattr(L, s) is not None


def test_abi_struct_sizes_match_header_comments():
    from h264decode_b200 import capi
    assert capi.NAL_DTYPE.itemsize == 32 and capi.NAL_EXT_DTYPE.itemsize == 24
    assert capi.FINAL_DTYPE.itemsize == 32 and ctypes.sizeof(capi.ScanSummary) == 48
    assert capi.lib().h264b_version() == 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from h264decode_b200 import capi
    with pytest.raises(capi.H264BError) as e:
        capi.Context(0)
    assert e.value.code == capi.E_NO_DEVICE


def test_product_never_touches_the_oracle():
    """nothing under h264decode_b200/ or include/ may reference oracle/ or harness/"""
    bad = []
    for base in ("h264decode_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath.split(os.sep) or "__pycache__" in dirpath:
                continue
            for f in files:
                if f.endswith((".so", ".o", ".pyc")):
                    continue
                t = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"(#include\s*[\"<][^\">]*oracle|import\s+oracle|from\s+oracle|liboracle|import\s+harness|from\s+harness)", t):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_generated_tables_match_reference_when_present():
    if not os.path.isdir("/root/reference/h264"):
        pytest.skip("reference not mounted (GPU box)")
    subprocess.check_call(["python", os.path.join(ROOT, "tools", "extract_tables.py"), "--check"])


def test_cut_byte_ranges_c_abi_matches_python_twin():
    """h264b_cut_byte_ranges is host-only (no context, no GPU): same ranges as sharding.cut_byte_ranges, bad arguments
    rejected"""
    import ctypes as C
    import numpy as np
    from h264decode_b200 import capi, sharding
    rng = np.random.default_rng(21)
    for t in range(60):
        n = int(rng.integers(0, 9000))
        s = rng.integers(0, 256, n, dtype=np.uint8)
        s[rng.random(n) < 0.5] = 0
        for pos in rng.integers(0, max(1, n - 4), max(1, n // 150)) if n >= 4 else []:
            s[pos:pos + 4] = [0, 0, 0, 1]
        for n_ranges in (1, 2, 3, 8, 13):
            assert capi.cut_byte_ranges(s, n_ranges) == sharding.cut_byte_ranges(s, n_ranges)
    lib = capi.lib()
    one = (C.c_uint64 * 1)()
    assert lib.h264b_cut_byte_ranges(None, 10, 1, one, one) != 0      # bytes announced, no buffer
    assert lib.h264b_cut_byte_ranges(None, 0, 0, one, one) != 0       # no ranges asked for
    assert lib.h264b_cut_byte_ranges(None, 0, 1, one, one) == 0 and one[0] == 0
