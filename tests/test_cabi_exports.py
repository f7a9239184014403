"""CPU: the C-ABI library loads and exports exactly the entry points include/cuco_b200.h declares;
argument errors are reported through the status/last_error convention without touching a GPU."""
import ctypes as C
import re
from pathlib import Path

from cucollections_b200 import _cabi

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "cuco_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cuco_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_cabi.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(native_lib):
    dll = C.CDLL(str(native_lib.path))
    for name in declared_symbols():
        assert hasattr(dll, name), name
    assert native_lib.flavour == "native"


def test_argument_errors_do_not_need_a_gpu(native_lib):
    handle = C.c_void_p()
    rc = native_lib.create(99, 100, 0.0, -1, -1, 0, 0, None, C.byref(handle))
    assert rc == 3 and b"kind" in native_lib.last_error()
    rc = native_lib.create(1, 100, 0.5, -1, -1, 1, -2, None, C.byref(handle))
    assert rc == 3  # erased-key constructor takes a capacity
    rc = native_lib.insert(None, None, None, 10, None, None)
    assert rc == 3
    assert native_lib.set_tuning(12, 0, 1, 0, 0, 1, 0) == 0
    assert native_lib.capacity(None) == -1


def test_product_path_has_no_cpu_fallback():
    """The package never imports the oracle, and a missing library is an error, not a fallback."""
    import cucollections_b200
    pkg = Path(cucollections_b200.__file__).parent
    for py in pkg.glob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, py
    try:
        _cabi.Library(pkg / "does_not_exist.so")
    except FileNotFoundError as e:
        assert "no CPU fallback" in str(e)
    else:
        raise AssertionError("missing library must raise")
