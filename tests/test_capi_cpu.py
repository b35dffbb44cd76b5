"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/metriko_b200.h
declares, and refuses loudly to compute without a B200 (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from metriko_b200 import capi


def test_library_exports_every_declared_symbol():
    names = capi.declared_symbols()
    assert len(names) >= 25
    L = C.CDLL(capi.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    # nothing mk_* is exported that the header does not declare
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (mk_[a-z0-9_]+)", out)))
    assert exported == names


def test_header_is_plain_c():
    src = open(capi.HEADER_PATH).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for bad in ("std::", "torch", "at::Tensor", "template", "class "):
        assert bad not in code
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", capi.HEADER_PATH], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_default_opts_and_no_cpu_fallback():
    o = capi.default_opts()
    assert o.tol == 1e-10 and o.max_iter == 100000 and o.precond == 1 and o.use_graph == 1
    if capi.lib().mk_device_count() == 0:
        with pytest.raises(capi.MkError) as e:
            capi.Context(0)
        assert e.value.status == -1  # MK_ERR_CUDA: the product path fails loudly without a device


def test_product_does_not_import_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for base, _, files in os.walk(os.path.join(root, "metriko_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                src = open(os.path.join(base, f), errors="ignore").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), os.path.join(base, f)
