"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol
include/sdplrp_b200.h declares, and refuses to run without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "sdplrp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdplrp_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(sp):
    lib = sp._lib.load()
    names = _header_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/sdplrp_b200.h but not exported"
    assert sorted(sp._lib.EXPORTED_SYMBOLS) == names, "ctypes binding and header disagree"


def test_version_and_error_strings(sp):
    lib = sp._lib.load()
    assert lib.sdplrp_version() >= 100
    assert b"no CUDA device" in lib.sdplrp_error_string(-6)


def test_no_cpu_fallback(sp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(sp.SdplrpError) as e:
        sp.Handle()
    assert e.value.code == sp._lib.ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sdplrplus.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "liboracle" not in text and "orc_" not in text, f


def test_every_option_key_is_documented_in_the_header():
    """sdplrp_set_option keys (csrc/api.cu) and the tuning-knob comment of include/sdplrp_b200.h must list the same names."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    keys = set(re.findall(r'k == "([a-z_0-9]+)"', open(os.path.join(root, "sdplrplus.jl_b200", "csrc", "api.cu")).read()))
    hdr = open(os.path.join(root, "include", "sdplrp_b200.h")).read()
    assert keys, "no option keys found"
    missing = sorted(k for k in keys if f'"{k}"' not in hdr)
    assert not missing, f"undocumented option keys: {missing}"
