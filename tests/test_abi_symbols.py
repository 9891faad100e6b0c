"""The C-ABI library loads (no GPU needed) and exports every function include/romis_gpu.h declares."""
import ctypes
import os
import re

from romis_b200 import abi
from romis_b200.api import EXPORTS, LIB_PATH

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "romis_gpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(romis_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from romis_b200 import build
    build.build()                      # nvcc cross-compiles sm_100a without a GPU
    assert os.path.exists(LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    names = declared_functions()
    assert len(names) >= 20
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/romis_gpu.h but not exported: {missing}"
    assert sorted(EXPORTS) == names, "romis_b200.api.EXPORTS out of sync with include/romis_gpu.h"
    lib.romis_abi_version.restype = ctypes.c_int
    assert lib.romis_abi_version() == 1


def test_struct_layouts_match_header():
    # sizes the C side static-asserts implicitly through its packing; a mismatch would corrupt every call
    assert ctypes.sizeof(abi.romis_vertex) == 32
    assert ctypes.sizeof(abi.romis_light) == 88
    assert ctypes.sizeof(abi.romis_features) == 64
    assert ctypes.sizeof(abi.romis_camera) == 36
    assert ctypes.sizeof(abi.romis_rng) == 16
    assert ctypes.sizeof(abi.romis_material) == 36


def test_product_never_imports_the_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "romis_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in text and "liboracle" not in text and "restir_oracle" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert not re.search(r'#include\s+"(\.\./)*oracle/', text), f
