"""The C-ABI library loads, exports every symbol include/cuba_b200.h declares, and refuses to compute
without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def _declared_functions():
    text = open(os.path.join(ROOT, "include", "cuba_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cuba_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg.library_path())
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libcuba_b200.so lacks %s" % n
    # the binding's list is the same set
    assert sorted(pkg.binding.exported_symbols()) == names


def test_cpp_api_symbols_present(pkg):
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", "-C", pkg.library_path()], capture_output=True, text=True).stdout
    assert "cuba::CudaBundleAdjustment::create()" in out
    assert "cuba::CudaBundleAdjustment::~CudaBundleAdjustment()" in out


def test_only_sm100a_code_is_embedded(pkg):
    import subprocess
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", pkg.library_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.CubaError, match="no CPU fallback"):
        pkg.Engine()


def test_product_does_not_touch_the_oracle():
    """the product package and its native sources never import / link anything under oracle/"""
    pdir = os.path.join(ROOT, "cuda-bundle-adjustment_b200")
    for dirpath, _, files in os.walk(pdir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "ba_oracle" not in text and "libcuba_ref" not in text, f
                assert not re.search(r"^\s*(import|from)\s+oracle", text, flags=re.M), f
