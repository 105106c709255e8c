"""The C-ABI library loads on a CPU-only box and exports every symbol include/bonej_thickness.h
declares; without a CUDA device creating a context fails loudly (no fallback)."""
import ctypes
import os
import re

import pytest

from bonej2_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "bonej_thickness.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(bjt_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def built():
    return _native.build()


def test_exports_every_declared_symbol(built):
    L = ctypes.CDLL(built)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in the header but not exported"


def test_binding_covers_header(built):
    assert set(_declared()) == set(_native.exported_symbols())


def test_no_device_is_an_error(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_native.ThicknessError) as e:
        _native.Context(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "bonej2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "lt_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


def test_product_knows_nothing_of_the_test_emulation():
    """the CPU emulation of CUDA, the JNI test double and the oracle live under tests/ and oracle/; nothing in the
    package, the headers, bench.py's product arm or the build of the product library refers to them"""
    pkg = os.path.join(ROOT, "bonej2_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                for word in ("cuda_emu", "kernels_emu", "emu_native", "libbonejthickness_emu", "jni_stub", "BJT_CUDA_EMU"):
                    assert word not in src, (f, word)
    from bonej2_b200 import _native
    assert os.path.basename(_native.SO_PATH) == "libbonejthickness.so" and os.path.dirname(_native.SO_PATH) == pkg
    assert not any("tests" in os.path.relpath(p, ROOT).split(os.sep) for p in _native.sources())


def test_paint_block_cols_is_a_host_function():
    """the column-block width rule needs no device (sharded runs query it to agree on one width)"""
    from bonej2_b200 import _native
    C = _native.Context
    assert C.paint_block_cols(1024, 1024, 1024) == 512          # 2^29 voxels per block
    assert C.paint_block_cols(1024, 1024, 128) == 4096          # not clipped to the width
    assert C.paint_block_cols(2048, 2048, 256) == 1024 and C.paint_block_cols(2048, 2048, 300) == 768
    assert C.paint_block_cols(64, 64, 64) % 128 == 0
    with pytest.raises(ValueError):
        C.paint_set_block_cols(100)                             # multiples of 32 only
    C.paint_set_block_cols(512)
    C.paint_set_block_cols(0)
