"""CPU suite: libhulo_gpu.so loads, exports every symbol include/hulo_gpu.h declares, and
fails loudly (no fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hulo_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hulo_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from sfmlocalization_b200 import _lib
    assert declared_symbols() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol():
    from sfmlocalization_b200 import _lib
    lib = _lib.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.hulo_version()


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "sfmlocalization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower(), os.path.join(dirpath, f)


def test_no_device_is_a_loud_failure():
    from sfmlocalization_b200 import _lib
    from sfmlocalization_b200.gpu import HuloError, HuloGpu
    lib = _lib.load()
    if lib.hulo_device_count() > 0:
        pytest.skip("a GPU is present; covered by the -m gpu suite")
    with pytest.raises(HuloError) as e:
        HuloGpu(0)
    assert e.value.status == _lib.ERR_CUDA
    assert "no CPU fallback" in str(e.value)


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/hulo_gpu.h must compile as C99 on its own."""
    import subprocess
    src = tmp_path / "hdr.c"
    src.write_text('#include "hulo_gpu.h"\nint main(void) { return hulo_device_count() < 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-c", str(src), "-o", str(tmp_path / "hdr.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
