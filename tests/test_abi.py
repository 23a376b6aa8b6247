"""The C-ABI library must load and export every entry point include/june_b200.h declares
(no compute calls here: there is no GPU in the build container)."""
import ctypes
import os
import re

from june_b200 import _lib

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "june_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jb_[a-z_0-9]+)\s*\(", text)))


def test_header_and_python_symbol_lists_agree():
    assert declared_symbols() == sorted("jb_" + n for n in _lib.ABI_SYMBOLS)


def test_cuda_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "libjune_b200.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} missing from libjune_b200.so"
    lib.jb_abi_version.restype = ctypes.c_int
    assert lib.jb_abi_version() == _lib.ABI_VERSION


def test_oracle_exports_the_same_abi():
    from helpers import oracle_lib
    lib = oracle_lib()
    for name in declared_symbols():
        assert hasattr(lib, "jo_" + name[3:]), f"jo_{name[3:]} missing from the oracle"


def test_product_package_never_references_the_oracle():
    pkg = os.path.join(ROOT, "june_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "june_oracle" not in text and "libjune_oracle" not in text, f
