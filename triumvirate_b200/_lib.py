"""Loader for the in-tree native libraries (no fallback: a missing or
unloadable library is an error, never a silent CPU path)."""
import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_LIBTRV = _HERE / "libtrv_b200.so"
_LIBTRVB = _HERE / "libtrvb.so"

_trv = None
_trvb = None


class NativeLibraryError(RuntimeError):
    pass


def trvb():
    """Device layer (C ABI of include/trvb.h)."""
    global _trvb
    if _trvb is None:
        if not _LIBTRVB.exists():
            raise NativeLibraryError(
                f"{_LIBTRVB} not built: run `python -c 'import __graft_entry__ "
                "as g; g.build()'` or `make -C triumvirate_b200`."
            )
        _trvb = C.CDLL(str(_LIBTRVB))
        _trvb.trvb_last_error.restype = C.c_char_p
        _trvb.trvb_version.restype = C.c_char_p
        _trvb.trvb_launch_count.restype = C.c_longlong
        _trvb.trvb_ctx_nmesh.restype = C.c_longlong
        _trvb.trvb_mesh_bytes.restype = C.c_size_t
        _trvb.trvb_ctx_stream.restype = C.c_void_p
        _trvb.trvb_cat_size.restype = C.c_longlong
    return _trvb


def trv():
    """Host C++ API with flat C entry points (src/capi.cpp)."""
    global _trv
    if _trv is None:
        trvb()
        if not _LIBTRV.exists():
            raise NativeLibraryError(f"{_LIBTRV} not built: run `make -C triumvirate_b200`.")
        _trv = C.CDLL(str(_LIBTRV))
        _trv.trv_last_error.restype = C.c_char_p
        _trv.trv_w3j.restype = C.c_double
        _trv.trv_coupling.restype = C.c_double
        _trv.trv_sjl_exact.restype = C.c_double
    return _trv
