"""TEST INFRASTRUCTURE ONLY -- builds the reference's OWN Cython binding layer
(``T/_threept.pyx``, ``_twopt.pyx``, ``_particles.pyx``, ``dataobjs.pyx``, ``parameters.pyx`` with
their ``.pxd`` files) UNMODIFIED against the headers and the shared library of
triumvirate_b200, as the drop-in proof of SURVEY.md section 8b: the extern
blocks ``T/_threept.pyx:27-95``, ``T/_particles.pxd:7-15``, ``T/dataobjs.pxd:16-128``
and ``T/parameters.pxd:8-86`` are the reference's binding contract.

The Cython sources are read from ``/root/reference`` into a temporary directory
(never into this repo); only the compiled extension modules are kept, in
``oracle/_ref/trvcy/`` (git-ignored; travels to the GPU box like libtrv_ref.so).
The package is called ``trvcy`` because ``import triumvirate`` itself needs
astropy, which is absent from the image.

A second package, ``trvcy_b200``, is built from the same sources after
``bindings/patch_bindings.py`` applied the Python-boundary fixes of SURVEY.md section 8f
rank 3 (memcpy LOS marshalling, ``except +`` on the estimator externs, raw-array catalogue
upload): the binding a maintainer would ship with the B200 build.

    python oracle/build_refcy.py          # needs /root/reference and a built libtrv_b200.so
"""
import shutil
import subprocess
import sys
import tempfile
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent
REF = Path("/root/reference/src/triumvirate")
OUT = HERE / "_ref" / "trvcy"
OUT_PATCHED = HERE / "_ref" / "trvcy_b200"
MODULES = ("parameters", "dataobjs", "_particles", "_threept", "_twopt")

SETUP = '''
from setuptools import setup, Extension
from Cython.Build import cythonize
import numpy as np
R = "{pkg}"
exts = [Extension("{name}." + m, ["{name}/" + m + ".pyx"], language="c++",
                  include_dirs=[np.get_include(), R + "/include/trv_compat", "{root}/include",
                                "{name}"],
                  library_dirs=[R], libraries=["trv_b200"],
                  extra_link_args=["-Wl,-rpath,$ORIGIN/../../../triumvirate_b200"],
                  extra_compile_args=["-std=c++17", "-w"])
        for m in {modules!r}]
# directives of the reference's own setup.py:1014-1018
setup(name="{name}", ext_modules=cythonize(
    exts, compiler_directives={{"language_level": "3", "c_string_encoding": "utf-8",
                               "embedsignature": True}},
    include_path=["{name}"], nthreads={nthreads}))
'''


def available(patched=False):
    out = OUT_PATCHED if patched else OUT
    return out.exists() and all(len(list(out.glob(f"{m}.*.so"))) == 1 for m in MODULES)


def build(force=False, patched=False):
    if available(patched) and not force:
        return True
    if not REF.exists() or not (ROOT / "triumvirate_b200" / "libtrv_b200.so").exists():
        return available(patched)
    name = "trvcy_b200" if patched else "trvcy"
    out = OUT_PATCHED if patched else OUT
    with tempfile.TemporaryDirectory() as tmp:
        pkg = Path(tmp) / name
        pkg.mkdir()
        if patched:
            sys.path.insert(0, str(ROOT / "bindings"))
            import patch_bindings
            patch_bindings.patch_tree(REF, pkg)
        else:
            for m in MODULES:
                shutil.copy(REF / f"{m}.pyx", pkg)
                if (REF / f"{m}.pxd").exists():
                    shutil.copy(REF / f"{m}.pxd", pkg)
        (pkg / "__init__.py").write_text("")
        (Path(tmp) / "setup.py").write_text(SETUP.format(
            pkg=ROOT / "triumvirate_b200", root=ROOT, modules=MODULES, nthreads=len(MODULES),
            name=name))
        env = dict(__import__("os").environ, CC="/usr/bin/gcc", CXX="/usr/bin/g++")
        subprocess.run([sys.executable, "setup.py", "-q", "build_ext", "--inplace",
                        "-j", str(len(MODULES))], cwd=tmp, check=True, env=env,
                       stdout=subprocess.DEVNULL)
        out.mkdir(parents=True, exist_ok=True)
        for so in out.glob("*.so"):
            so.unlink()
        for so in pkg.glob("*.so"):
            shutil.copy(so, out)
        (out / "__init__.py").write_text(
            "# compiled from the reference's Cython sources"
            + (" patched by bindings/patch_bindings.py" if patched else " (unmodified)")
            + " by oracle/build_refcy.py\n")
    return available(patched)


if __name__ == "__main__":
    ok = build(force=True) and build(force=True, patched=True)
    print("oracle/_ref/trvcy, oracle/_ref/trvcy_b200:", "built" if ok else "NOT built")
    sys.exit(0 if ok else 1)
