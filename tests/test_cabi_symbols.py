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


def test_knob_constants_match_header_and_library(lib_path):
    """MSDA_KNOB_* of the header == KNOB_* of the ctypes face; msda_set_knob (host-only state, no GPU needed) accepts exactly
    those indices, reports the documented defaults and restores what it is given."""
    from uninext_b200 import _cabi
    text = open(HEADER).read()
    hdr = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+MSDA_KNOB_(\w+)\s+(\d+)\b", text)}
    count = hdr.pop("COUNT")
    assert sorted(hdr.values()) == list(range(count))
    for name, idx in hdr.items():
        assert getattr(_cabi, "KNOB_" + name) == idx, name
    lib = _cabi.load()
    query = -1000000                                                        # MSDA_KNOB_QUERY
    assert lib.msda_set_knob(count, query) < 0 and lib.msda_set_knob(-1, query) < 0          # MSDA_E_BADARG
    if "MSDA_ZERO_FILL" not in os.environ:
        assert lib.msda_set_knob(_cabi.KNOB_ZERO_FILL, query) == 2         # default: fill kernel as PDL primary
    if "MSDA_SLAB" not in os.environ:
        assert lib.msda_set_knob(_cabi.KNOB_SLAB, query) == -1             # auto = tiled kernels
    for idx in range(count):
        was = lib.msda_set_knob(idx, query)
        assert lib.msda_set_knob(idx, 7) == was and lib.msda_set_knob(idx, query) == 7
        assert lib.msda_set_knob(idx, was) == 7 and lib.msda_set_knob(idx, query) == was
