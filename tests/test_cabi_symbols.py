"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads without a
GPU (static cudart), and exports every symbol include/oscillink_b200.h declares.  No compute
calls are made here."""
import ctypes
import os
import re

from oscillink_b200 import _cabi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "oscillink_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(osc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported():
    lib = ctypes.CDLL(build.build())
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_ctypes_prototypes_cover_header():
    assert sorted(_cabi.PROTOTYPES) == _declared_symbols()


def test_library_loads_and_reports_version():
    lib = _cabi.load()
    assert lib.osc_abi_version() == 1
    assert isinstance(lib.osc_last_error(), bytes)


def test_invalid_argument_maps_to_value_error_without_gpu():
    lib = _cabi.load()
    need = ctypes.c_size_t(0)
    rc = lib.osc_knn_build_workspace(1, 10, 0, 3, 0, ctypes.byref(need))  # D = 0
    assert rc == _cabi.ERR_INVALID
    try:
        _cabi.check(rc)
    except ValueError as e:
        assert "knn_build_workspace" in str(e)
    else:
        raise AssertionError("expected ValueError")


def test_struct_layouts_match_header_sizes():
    # sizes implied by the C declarations (LP64)
    assert ctypes.sizeof(_cabi.Graph) == 8 + 8 + 4 + 4 + 5 * 8
    assert ctypes.sizeof(_cabi.Chain) == 8 + 6 * 8
    assert ctypes.sizeof(_cabi.Params) == 24
    assert ctypes.sizeof(_cabi.PcgDims) == 32
    assert ctypes.sizeof(_cabi.BatchedArgs) == 8 * 8 + 16 + 8 + 16 + 8 + 8  # + unresolved*
