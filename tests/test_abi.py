"""The C-ABI: the CUDA library loads without a GPU and exports every function include/cczero_b200.h declares; the
ctypes table covers the header; the product loader refuses a missing library (no CPU fallback)."""
import ctypes
import importlib
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "cczero_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cz_[a-z0-9_]+)\s*\(", src)))


def test_cuda_library_exports_header():
    build = importlib.import_module("chinesechess-alphazero_b200.build")
    path = build.build_cuda()
    dll = ctypes.CDLL(path)
    names = header_functions()
    assert len(names) >= 35
    for n in names:
        assert hasattr(dll, n), f"{n} declared in the header but not exported"
    assert dll.cz_build_is_cuda() == 1


def test_ctypes_table_matches_header():
    from cczero_b200.lib import _SIGS
    assert sorted(_SIGS) == header_functions()


def test_host_only_entry_points():
    from cczero_b200.lib import CzLib, CUDA_LIB_PATH
    lib = CzLib(CUDA_LIB_PATH)
    assert not lib.missing
    labels = ctypes.create_string_buffer(2086 * 4)
    lut = (ctypes.c_int16 * 8100)()
    lib.call("cz_action_labels", ctypes.cast(labels, ctypes.c_void_p), ctypes.cast(lut, ctypes.c_void_p))
    raw = labels.raw.decode()
    assert raw[:4] == "0010" and raw[-4:] == "8765"
    assert sorted(v for v in lut if v >= 0) == list(range(2086))
    with pytest.raises(Exception):
        lib.call("cz_action_labels", None, None)
    assert b"NULL" in lib.raw("cz_last_error")()


def test_no_cpu_fallback(tmp_path):
    from cczero_b200.lib import CzError, CzLib
    with pytest.raises(CzError):
        CzLib(str(tmp_path / "libcczero_b200.so"))
    import cczero_b200.lib as L
    src = open(L.__file__).read()
    assert "oracle" not in src


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "chinesechess-alphazero_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            txt = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", txt, flags=re.M), fn
            assert "libcz_emul" not in txt or fn == "build.py", fn   # only the build recipe knows the emulator library
