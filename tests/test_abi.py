"""The C-ABI libraries load on a CPU-only box and export every symbol the headers declare; the CUDA
library refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from pingpongpro_b200 import _build, api

INC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")


def declared(header):
    text = open(os.path.join(INC, header)).read()
    return sorted(set(re.findall(r"PPP_API\s+[\w\s\*]+?\b(ppp_\w+)\s*\(", text)))


def test_cuda_library_exports_every_declared_symbol():
    names = declared("ppp_b200.h")
    assert len(names) >= 20
    L = ctypes.CDLL(_build.build_cuda())
    for n in names:
        assert hasattr(L, n), n


def test_host_library_exports_every_declared_symbol():
    names = declared("ppp_host.h")
    assert len(names) >= 10
    L = ctypes.CDLL(_build.build_host())
    for n in names:
        assert hasattr(L, n), n


def test_struct_layouts_match_header_sizes():
    assert api.LOCUS_DTYPE.itemsize == 20 and api.PAIR_DTYPE.itemsize == 40 and api.INTERVAL_DTYPE.itemsize == 16
    assert api.TP_RESULT_DTYPE.itemsize == 160 and ctypes.sizeof(api.Timings) == 40 and ctypes.sizeof(api.Config) == 56


def test_no_cpu_fallback_without_device():
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(api.PppError) as e:
        api.Context([1000, 2000])
    assert e.value.code == api.ERR_CUDA and "no CPU fallback" in str(e.value)


def test_argument_validation_happens_before_cuda():
    for kw in (dict(max_overlap=40), dict(min_overlap=12), dict(mode=7)):
        with pytest.raises(api.PppError) as e:
            api.Context([1000], **kw)
        assert e.value.code == api.ERR_ARG
    with pytest.raises(api.PppError):
        api.Context([])


def test_product_code_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may use oracle/ (the checker is not the product)."""
    pkg = os.path.dirname(api.__file__)
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h")) and f != "_build.py":
                text = open(os.path.join(root, f), errors="replace").read()
                assert "ppp_oracle" not in text and "oracle_tools" not in text and "oracle/" not in text.replace("oracle/seqan_shim", ""), f
