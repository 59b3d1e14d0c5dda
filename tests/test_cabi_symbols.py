"""CPU test: the C-ABI library loads and exports every symbol include/msda_b200.h declares, and the ctypes table in
uninext_b200/_cabi.py mirrors the header (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "msda_b200.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|uint64_t|const char \*)\s*(msda_\w+)\s*\(([^;{]*)\)\s*;", text):
        args = [a.strip() for a in m.group(2).split(",") if a.strip() and a.strip() != "void"]
        decls[m.group(1)] = len(args)
    return decls


@pytest.fixture(scope="module")
def lib_path():
    from uninext_b200 import build
    return build.build()


def test_header_declares_the_expected_entry_points():
    d = _declared()
    for name in ("msda_forward_f32", "msda_forward_f64", "msda_forward_bf16", "msda_backward_f32", "msda_backward_f64",
                 "msda_backward_bf16", "msda_abi_version", "msda_strerror", "msda_uses_fast_path", "msda_launch_count"):
        assert name in d, name


def test_library_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in msda_b200.h but not exported by {lib_path}"


def test_ctypes_table_matches_header(lib_path):
    from uninext_b200 import _cabi
    decl = _declared()
    assert set(decl) == set(_cabi.SIGNATURES)
    for name, nargs in decl.items():
        assert len(_cabi.SIGNATURES[name][1]) == nargs, name
    lib = _cabi.load()
    assert lib.msda_abi_version() == _cabi.ABI_VERSION
    assert b"bad argument" in lib.msda_strerror(-1)
    assert lib.msda_uses_fast_path(4, 32, 4, 4) == 1 and lib.msda_uses_fast_path(8, 32, 4, 4) == 0


def test_missing_library_fails_loudly(tmp_path):
    from uninext_b200 import _cabi
    with pytest.raises(_cabi.MSDALibraryError, match="no CPU / PyTorch fallback"):
        _cabi.load(str(tmp_path / "nope.so"))
