"""CPU suite, part 2: the C-ABI library loads, exports every symbol include/mpgpu.h declares,
and refuses to compute without a CUDA device (no CPU fallback).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mpgpu.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(mpgpu_[a-z0-9_]+)\s*\(", text))
    names -= {"mpgpu_rng_fn"}
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from mpboot_b200 import engine
    L = engine.lib()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "libmpgpu.so does not export %s" % s


def test_no_cpu_fallback():
    from mpboot_b200 import engine
    if engine.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(engine.MpGpuError, match="no CPU fallback"):
        engine.Engine()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "mpboot_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "mp_oracle" not in src, f
