"""Python-boundary fixes for the reference's Cython layer,
shipped as CODE: this script rewrites the reference's own ``_threept.pyx``,
``_twopt.pyx``, ``_particles.pxd`` and ``_particles.pyx`` -- read from the reference
tree at build time, never stored in this repository -- with three rule-based edits:

1. LOS marshalling (``T/_threept.pyx:138-152`` and its clones, ``T/_twopt.pyx``): the
   per-particle Python loop ``for pid, (los_x, los_y, los_z) in enumerate(los): ...``
   that fills the malloc'ed ``LineOfSight`` array costs minutes at 5e7 randoms; the
   ``(N, 3)`` C-contiguous float64 array IS an array of ``LineOfSight`` (three doubles,
   ``I/dataobjs.hpp:186-188``), so one ``memcpy`` replaces the loop.
2. ``except +`` on every ``compute_*`` / ``calc_*`` extern (``T/_threept.pyx:50-95``):
   a C++ exception from the estimators (e.g. ``trv::sys::DeviceError`` when no GPU is
   usable) becomes a Python ``RuntimeError`` instead of terminating the interpreter.
3. Catalogue upload (``T/_particles.pxd:11-14``, ``T/_particles.pyx:30``): the six
   ``std::vector<double>`` taken BY VALUE are built element by element from the numpy
   arrays by Cython and then copied again by C++; the patched binding hands the array
   pointers to ``trv::ParticleCatalogue::load_particle_arrays`` (one pass, no temporaries).

    python bindings/patch_bindings.py <reference package dir> <output dir>

``oracle/build_refcy.py --patched`` applies it and builds the result against
``libtrv_b200.so`` as the package ``trvcy_b200`` (tests/test_dropin_cython.py).
"""
import re
import sys
from pathlib import Path

LOS_LOOP = re.compile(
    r"^(?P<ind>[ \t]*)for pid, \(los_x, los_y, los_z\) in enumerate\((?P<arr>\w+)\):\n"
    r"(?:(?P=ind)[ \t]+(?P<dst>\w+)\[pid\]\.pos\[[012]\] = los_[xyz]\n){3}", re.M)


def patch_estimator_module(text):
    """Rules 1 and 2 on ``_threept.pyx`` / ``_twopt.pyx``."""
    n_los = 0

    def los_memcpy(m):
        nonlocal n_los
        n_los += 1
        ind, arr = m.group("ind"), m.group("arr")
        dst = re.search(r"(\w+)\[pid\]", m.group(0)).group(1)
        return (f"{ind}if {arr}.shape[0] > 0:\n"
                f"{ind}    memcpy({dst}, &{arr}[0, 0], {arr}.shape[0] * sizeof(LineOfSight))\n")

    text = LOS_LOOP.sub(los_memcpy, text)
    if n_los:
        text = text.replace("from libc.stdlib cimport free, malloc",
                            "from libc.stdlib cimport free, malloc\nfrom libc.string cimport memcpy", 1)

    # `except +` on extern declarations that lack it: a declaration ends with the line
    # holding the closing parenthesis of its argument list inside a `cdef extern` block
    out, in_extern, depth, n_exc = [], False, 0, 0
    for line in text.split("\n"):
        stripped = line.strip()
        if line.startswith("cdef extern from"):
            in_extern = True
        elif in_extern and line and not line[0].isspace() and not line.startswith("#"):
            in_extern = False
        if in_extern and not stripped.startswith("#"):
            opens, closes = line.count("("), line.count(")")
            was_open = depth > 0
            depth += opens - closes
            if (was_open or opens) and depth == 0 and stripped.endswith(")"):
                line = line.rstrip() + " except +"
                n_exc += 1
        out.append(line)
    return "\n".join(out), n_los, n_exc


def patch_particles_pxd(text):
    """Rule 3, declaration side: add the raw-array loader to the extern class."""
    anchor = re.search(r"^(?P<ind>[ \t]+)int load_particle_data\(", text, re.M)
    if not anchor:
        raise RuntimeError("load_particle_data declaration not found")
    ind = anchor.group("ind")
    decl = (f"{ind}int load_particle_arrays(\n"
            f"{ind}    int n, const double* x, const double* y, const double* z,\n"
            f"{ind}    const double* nz, const double* ws, const double* wc\n"
            f"{ind}) except +\n\n")
    return text[:anchor.start()] + decl + text[anchor.start():]


def patch_particles_pyx(text):
    """Rule 3, call side."""
    pat = re.compile(r"^(?P<ind>[ \t]+)self\.thisptr\.load_particle_data\(x, y, z, nz, ws, wc\)\n", re.M)
    m = pat.search(text)
    if not m:
        raise RuntimeError("load_particle_data call not found")
    ind = m.group("ind")
    call = (f"{ind}if not (len(x) == len(y) == len(z) == len(nz) == len(ws) == len(wc)):\n"
            f"{ind}    raise ValueError('Inconsistent particle data dimensions.')\n"
            f"{ind}if len(x) == 0:\n"
            f"{ind}    raise ValueError('Particle data are empty.')\n"
            f"{ind}self.thisptr.load_particle_arrays(\n"
            f"{ind}    <int>len(x), &x[0], &y[0], &z[0], &nz[0], &ws[0], &wc[0]\n"
            f"{ind})\n")
    return text[:m.start()] + call + text[m.end():]


def patch_tree(src, dst):
    """Copy the binding sources of `src` into `dst`, patched.  Returns a summary."""
    src, dst = Path(src), Path(dst)
    dst.mkdir(parents=True, exist_ok=True)
    summary = {}
    for name in ("parameters", "dataobjs", "_particles", "_threept", "_twopt"):
        for ext in (".pyx", ".pxd"):
            f = src / (name + ext)
            if not f.exists():
                continue
            text = f.read_text()
            if name in ("_threept", "_twopt") and ext == ".pyx":
                text, n_los, n_exc = patch_estimator_module(text)
                summary[f.name] = {"los_loops_replaced": n_los, "except_plus_added": n_exc}
            elif name == "_particles" and ext == ".pxd":
                text = patch_particles_pxd(text)
                summary[f.name] = {"raw_array_loader_declared": True}
            elif name == "_particles" and ext == ".pyx":
                text = patch_particles_pyx(text)
                summary[f.name] = {"raw_array_loader_called": True}
            (dst / f.name).write_text(text)
    return summary


if __name__ == "__main__":
    import json
    print(json.dumps(patch_tree(sys.argv[1], sys.argv[2]), indent=1))
