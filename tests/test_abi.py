"""CPU tests: both in-tree shared libraries load without a GPU and export every
symbol their C headers (include/*.h) declare; the product path fails loudly --
never falls back to a CPU implementation -- when no device is usable."""
import ctypes as C
import re
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_functions(header):
    text = (ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    # identifiers directly followed by "(" at declaration level
    names = set(re.findall(r"\b(trvb?_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


@pytest.mark.parametrize("header,loader", [("trvb.h", "trvb"), ("trv_capi.h", "trv")])
def test_library_exports_every_declared_symbol(header, loader):
    from triumvirate_b200 import _lib
    lib = getattr(_lib, loader)()
    names = _declared_functions(header)
    assert len(names) > 15, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"{header}: not exported: {missing}"


def test_headers_compile_as_c():
    """The boundary is a C ABI: plain C must be able to include the headers."""
    import subprocess, tempfile
    src = '#include "trvb.h"\n#include "trv_capi.h"\nint main(void){return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        f = Path(d) / "t.c"
        f.write_text(src)
        subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only",
                        f"-I{ROOT / 'include'}", str(f)], check=True)


def test_every_entry_point_cites_the_reference():
    for header in ("trvb.h", "trv_capi.h"):
        text = (ROOT / "include" / header).read_text()
        assert len(re.findall(r"[ST]/[a-z_]+\.(?:cpp|pyx|pxd|py):\d+", text)) >= 15, header


def test_no_cpu_fallback_without_a_device():
    from triumvirate_b200 import _lib, core
    if _lib.trvb().trvb_device_count() > 0:
        pytest.skip("a CUDA device is visible")
    assert core.gpu_count() == 0
    pos = np.random.default_rng(0).uniform(0., 100., size=(3, 100))
    with pytest.raises(core.TriumvirateError, match="no CPU fallback"):
        core.threept("bispec", "sim", pos, 100., 16, "tsc", (0, 0, 0), "diag",
                     (0.1, 0.5), 3, 1.)
    with pytest.raises(core.TriumvirateError):
        core.mesh(pos, 100., 16, "tsc")
    # two-point estimators and the array fast path likewise
    with pytest.raises(core.TriumvirateError, match="no CPU fallback"):
        core.twopt("powspec", "sim", 100., 16, "tsc", 0, (0.1, 0.5), 3, 1., pos_d=pos)
    x = np.ascontiguousarray(pos)
    with pytest.raises(core.TriumvirateError, match="no CPU fallback"):
        core.threept_box_arrays("bispec", 100, x[0].ctypes.data, x[1].ctypes.data,
                                x[2].ctypes.data, False, 100., 16, "tsc", (0, 0, 0), "diag",
                                (0.1, 0.5), 3)
    # host-only helpers keep working without a device
    own = core.partition_owners("full", (0, 0), 6, 4)
    assert len(own) == 21 and set(own) == {0, 1, 2, 3}
    assert abs(core.norm_particles_2pt(pos, np.full(100, 1.e-4)) - 1. / (100 * 1.e-4)) < 1e-9
    # the device layer itself refuses to create a context
    ctx = C.c_void_p()
    st = _lib.trvb().trvb_ctx_create(C.byref(ctx), 0, (C.c_int * 3)(8, 8, 8),
                                     (C.c_double * 3)(1., 1., 1.), 2)
    assert st != 0 and b"no CPU fallback" in _lib.trvb().trvb_last_error()


def test_product_package_never_imports_the_oracle():
    """Only tests/, smoke() and bench.py's baseline legs may touch oracle/."""
    for p in (ROOT / "triumvirate_b200").rglob("*"):
        if p.suffix in (".py", ".cpp", ".hpp", ".cu", ".cuh", ".h") or p.name == "Makefile":
            text = p.read_text()
            assert "oracle" not in text.lower() or "oracle" not in re.sub(
                r"(#|//)[^\n]*", "", text).lower(), f"{p} references oracle/"
    before = set(sys.modules)
    import triumvirate_b200  # noqa: F401
    from triumvirate_b200 import core  # noqa: F401
    assert not any(m == "oracle" or m.startswith("oracle.") for m in set(sys.modules) - before)
