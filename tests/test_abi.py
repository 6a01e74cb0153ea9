"""The C-ABI library loads and exports every symbol include/cassiopee_b200.h
declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "cassiopee_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cx_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_header_symbols():
    import __graft_entry__ as g
    so = os.path.join(ROOT, "cassiopee_b200", "csrc", "libcassiopee_b200.so")
    if not os.path.exists(so):
        g.build()
    L = ctypes.CDLL(so)
    syms = _header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(L, s), s


def test_binding_table_matches_header():
    from cassiopee_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _header_symbols()


def test_no_device_fails_loudly():
    """Without a CUDA device the product raises; it never falls back to a CPU path."""
    import pytest
    from cassiopee_b200 import _lib
    n = ctypes.c_int(0)
    rc = _lib.lib().cx_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(Exception):
        _lib.Context(0)


def test_entry_scripts_compile():
    """bench.py and its legs, __graft_entry__.py, the workers and the tools byte-compile (a repeated keyword argument is
    a SyntaxError at compile time that neither an import of the package nor the CPU tests would show)."""
    import glob
    import py_compile
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    files = glob.glob(os.path.join(root, "bench*.py")) + [os.path.join(root, "__graft_entry__.py")] + \
        glob.glob(os.path.join(root, "tests", "*.py")) + glob.glob(os.path.join(root, "tools", "*.py")) + \
        glob.glob(os.path.join(root, "cassiopee_b200", "**", "*.py"), recursive=True)
    assert len(files) > 20
    for f in files:
        py_compile.compile(f, doraise=True)
