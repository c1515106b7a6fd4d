"""TEST INFRASTRUCTURE ONLY -- ctypes front end to ``oracle/_ref/libtrv_ref.so``.

``libtrv_ref.so`` is the reference's own C++ (``/root/reference/src/triumvirate/
src/*.cpp``) compiled UNMODIFIED by ``oracle/Makefile`` against the local
FFTW3/GSL shim, plus the array-marshalling driver ``oracle/ref_driver.cpp``.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import this module; the product package
``triumvirate_b200`` never does.

The catalogue pre-processing helpers restate the reference's Python-side
semantics (``catalogue.py`` cannot be imported here: astropy is absent):

* :func:`periodise`  -- ``T/catalogue.py:648-676``
* :func:`centre`     -- ``T/catalogue.py:490-545``
* :func:`compute_los`-- ``T/catalogue.py:437-458``
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_ref" / "libtrv_ref.so"
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def available():
    return _LIB_PATH.exists()


def build(ref="/root/reference"):
    """Build ``oracle/_ref/libtrv_ref.so`` when the reference tree is present."""
    import subprocess
    if not Path(ref).exists():
        return available()
    subprocess.run(
        ["make", "-C", str(_HERE), f"REF={ref}", "-j8"],
        check=True, stdout=subprocess.DEVNULL,
    )
    return available()


def lib():
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f"{_LIB_PATH} is missing: run `make -C oracle` where "
                "/root/reference is mounted."
            )
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.trvref_last_error.restype = C.c_char_p
        _lib.trvref_w3j.restype = C.c_double
        _lib.trvref_coupling.restype = C.c_double
    return _lib


def _d(a):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _check(status):
    if status != 0:
        raise RuntimeError(lib().trvref_last_error().decode())


def set_num_threads(n):
    lib().trvref_set_num_threads(int(n))


def num_threads():
    return lib().trvref_num_threads()


# -- catalogue pre-processing (restated Python-side semantics) ------------

def periodise(pos, boxsize):
    """``(x + L/2 - (max + min)/2) % L`` per axis (T/catalogue.py:668-674)."""
    pos = np.array(pos, dtype=np.float64, copy=True)
    boxsize = np.broadcast_to(np.asarray(boxsize, dtype=np.float64), (3,))
    for ax in range(3):
        lo, hi = pos[ax].min(), pos[ax].max()
        pos[ax] = (pos[ax] + boxsize[ax] / 2. - (hi + lo) / 2.) % boxsize[ax]
    return pos


def centre(pos, pos_ref, boxsize):
    """Shift both catalogues so the reference's extent mid-point sits at the
    box centre (T/catalogue.py:535-545)."""
    boxsize = np.broadcast_to(np.asarray(boxsize, dtype=np.float64), (3,))
    pos = np.array(pos, dtype=np.float64, copy=True)
    pos_ref = np.array(pos_ref, dtype=np.float64, copy=True)
    origin = np.array([
        np.mean([pos_ref[ax].min(), pos_ref[ax].max()]) - boxsize[ax] / 2.
        for ax in range(3)
    ])
    for ax in range(3):
        pos[ax] -= origin[ax]
        pos_ref[ax] -= origin[ax]
    return pos, pos_ref


def compute_los(pos):
    """``pos/|pos|`` with zero-norm guard (T/catalogue.py:437-458); (N, 3)."""
    pos = np.asarray(pos, dtype=np.float64)
    norm = np.sqrt(pos[0]**2 + pos[1]**2 + pos[2]**2)
    norm[norm == 0.] = 1.
    return np.ascontiguousarray(
        np.transpose([pos[0] / norm, pos[1] / norm, pos[2] / norm])
    )


# -- reference entry points -------------------------------------------------

def threept(stat, catalogue_type, pos_d, boxsize, ngrid, assignment, degrees,
            form, bin_range, num_bins, norm_factor, idx_bin=0, binning="lin",
            nz_d=None, ws_d=None, wc_d=None, los_d=None,
            pos_r=None, nz_r=None, ws_r=None, wc_r=None, los_r=None,
            interlace=False, verbose=60):
    """Run the reference three-point estimator (S/threept.cpp:248-2619).

    Positions are (3, N) arrays ALREADY aligned in the box (see
    :func:`periodise`, :func:`centre`).  Returns a dict of numpy arrays.
    """
    L = lib()
    boxsize = np.broadcast_to(np.asarray(boxsize, dtype=np.float64), (3,)).copy()
    ngrid = np.broadcast_to(np.asarray(ngrid, dtype=np.int32), (3,)).copy()
    pos_d = np.asarray(pos_d, dtype=np.float64)
    nd = pos_d.shape[1]
    keep = []

    def arr(a):
        a_, p = _d(a)
        keep.append(a_)
        return p

    args_d = [arr(pos_d[0]), arr(pos_d[1]), arr(pos_d[2]),
              arr(nz_d), arr(ws_d), arr(wc_d), arr(los_d)]
    if pos_r is not None:
        pos_r = np.asarray(pos_r, dtype=np.float64)
        nr = pos_r.shape[1]
        args_r = [arr(pos_r[0]), arr(pos_r[1]), arr(pos_r[2]),
                  arr(nz_r), arr(ws_r), arr(wc_r), arr(los_r)]
    else:
        nr = 0
        args_r = [None] * 7
    nb = int(num_bins)
    cap = max(nb * nb, nb) + 8
    dim = C.c_int(0)
    c1b = np.zeros(cap); c2b = np.zeros(cap)
    c1e = np.zeros(cap); c2e = np.zeros(cap)
    n1 = np.zeros(cap, dtype=np.int32); n2 = np.zeros(cap, dtype=np.int32)
    raw = np.zeros(2 * cap); shot = np.zeros(2 * cap)
    elapsed = C.c_double(0.)
    status = L.trvref_threept(
        stat.encode(), catalogue_type.encode(),
        C.c_int(nd), *args_d, C.c_int(nr), *args_r,
        boxsize.ctypes.data_as(_dp), ngrid.ctypes.data_as(_ip),
        assignment.encode(),
        C.c_int(degrees[0]), C.c_int(degrees[1]), C.c_int(degrees[2]),
        form.encode(), C.c_int(idx_bin or 0), binning.encode(),
        C.c_double(bin_range[0]), C.c_double(bin_range[1]), C.c_int(nb),
        C.c_int(1 if interlace else 0), C.c_double(norm_factor),
        C.c_int(verbose), C.byref(dim),
        c1b.ctypes.data_as(_dp), c2b.ctypes.data_as(_dp),
        c1e.ctypes.data_as(_dp), c2e.ctypes.data_as(_dp),
        n1.ctypes.data_as(_ip), n2.ctypes.data_as(_ip),
        raw.ctypes.data_as(_dp), shot.ctypes.data_as(_dp), C.byref(elapsed),
    )
    _check(status)
    n = dim.value
    raw_c = raw[0:2*n:2] + 1j * raw[1:2*n:2]
    shot_c = shot[0:2*n:2] + 1j * shot[1:2*n:2]
    if stat == "bispec":
        names = ("k1_bin", "k2_bin", "k1_eff", "k2_eff", "nmodes_1",
                 "nmodes_2", "bk_raw", "bk_shot")
    else:
        names = ("r1_bin", "r2_bin", "r1_eff", "r2_eff", "npairs_1",
                 "npairs_2", "zeta_raw", "zeta_shot")
    vals = (c1b[:n].copy(), c2b[:n].copy(), c1e[:n].copy(), c2e[:n].copy(),
            n1[:n].copy(), n2[:n].copy(), raw_c, shot_c)
    out = dict(zip(names, vals))
    out["elapsed_s"] = elapsed.value
    return out



def threept_window(pos_r, boxsize, ngrid, assignment, degrees, form, bin_range,
                   num_bins, norm_factor, los_r, alpha=1., idx_bin=0, binning="lin",
                   nz_r=None, ws_r=None, wc_r=None, wide_angle=False,
                   wa_orders=(0, 0), verbose=60):
    """Run the reference 3PCF window estimator (S/threept.cpp:2621-3077)."""
    L = lib()
    boxsize = np.broadcast_to(np.asarray(boxsize, dtype=np.float64), (3,)).copy()
    ngrid = np.broadcast_to(np.asarray(ngrid, dtype=np.int32), (3,)).copy()
    pos_r = np.asarray(pos_r, dtype=np.float64)
    keep = [np.ascontiguousarray(a, dtype=np.float64) if a is not None else None
            for a in (pos_r[0], pos_r[1], pos_r[2], nz_r, ws_r, wc_r, los_r)]
    ptrs = [a.ctypes.data_as(_dp) if a is not None else None for a in keep]
    nb = int(num_bins)
    cap = max(nb * nb, nb) + 8
    dim = C.c_int(0)
    c1b = np.zeros(cap); c2b = np.zeros(cap)
    c1e = np.zeros(cap); c2e = np.zeros(cap)
    n1 = np.zeros(cap, dtype=np.int32); n2 = np.zeros(cap, dtype=np.int32)
    raw = np.zeros(2 * cap); shot = np.zeros(2 * cap)
    status = L.trvref_threept_window(
        C.c_int(pos_r.shape[1]), *ptrs,
        boxsize.ctypes.data_as(_dp), ngrid.ctypes.data_as(_ip), assignment.encode(),
        C.c_int(degrees[0]), C.c_int(degrees[1]), C.c_int(degrees[2]),
        C.c_int(wa_orders[0]), C.c_int(wa_orders[1]),
        form.encode(), C.c_int(idx_bin or 0), binning.encode(),
        C.c_double(bin_range[0]), C.c_double(bin_range[1]), C.c_int(nb),
        C.c_double(alpha), C.c_double(norm_factor), C.c_int(1 if wide_angle else 0),
        C.c_int(verbose),
        C.byref(dim),
        c1b.ctypes.data_as(_dp), c2b.ctypes.data_as(_dp),
        c1e.ctypes.data_as(_dp), c2e.ctypes.data_as(_dp),
        n1.ctypes.data_as(_ip), n2.ctypes.data_as(_ip),
        raw.ctypes.data_as(_dp), shot.ctypes.data_as(_dp),
    )
    _check(status)
    n = dim.value
    names = ("r1_bin", "r2_bin", "r1_eff", "r2_eff", "npairs_1", "npairs_2",
             "zeta_raw", "zeta_shot")
    vals = (c1b[:n].copy(), c2b[:n].copy(), c1e[:n].copy(), c2e[:n].copy(),
            n1[:n].copy(), n2[:n].copy(),
            raw[0:2*n:2] + 1j * raw[1:2*n:2], shot[0:2*n:2] + 1j * shot[1:2*n:2])
    return dict(zip(names, vals))

def twopt(stat, catalogue_type, boxsize, ngrid, assignment, degree, bin_range, num_bins,
          norm_factor, pos_d=None, nz_d=None, ws_d=None, wc_d=None, los_d=None,
          pos_r=None, nz_r=None, ws_r=None, wc_r=None, los_r=None,
          interlace=False, binning="lin", alpha=1., verbose=60):
    """Run the reference two-point estimator (S/twopt.cpp:388-901); same
    arguments and result names as ``triumvirate_b200.core.twopt``."""
    L = lib()
    boxsize = np.broadcast_to(np.asarray(boxsize, dtype=np.float64), (3,)).copy()
    ngrid = np.broadcast_to(np.asarray(ngrid, dtype=np.int32), (3,)).copy()
    keep = []

    def arr(a):
        a_, p = _d(a)
        keep.append(a_)
        return p

    def cat(pos, nz, ws, wc, los):
        if pos is None:
            return 0, [None] * 7
        pos = np.asarray(pos, dtype=np.float64)
        return pos.shape[1], [arr(pos[0]), arr(pos[1]), arr(pos[2]), arr(nz), arr(ws), arr(wc),
                              arr(los)]

    nd, args_d = cat(pos_d, nz_d, ws_d, wc_d, los_d)
    nr, args_r = cat(pos_r, nz_r, ws_r, wc_r, los_r)
    nb = int(num_bins)
    dim = C.c_int(0)
    cb = np.zeros(nb + 8); ce = np.zeros(nb + 8)
    cnt = np.zeros(nb + 8, dtype=np.int32)
    raw = np.zeros(2 * (nb + 8)); shot = np.zeros(2 * (nb + 8))
    elapsed = C.c_double(0.)
    if isinstance(interlace, str):
        interlace = interlace == "true"
    _check(L.trvref_twopt(
        stat.encode(), catalogue_type.encode(),
        C.c_int(nd), *args_d, C.c_int(nr), *args_r,
        boxsize.ctypes.data_as(_dp), ngrid.ctypes.data_as(_ip),
        assignment.encode(), C.c_int(1 if interlace else 0), C.c_int(degree), binning.encode(),
        C.c_double(bin_range[0]), C.c_double(bin_range[1]), C.c_int(nb),
        C.c_double(alpha), C.c_double(norm_factor), C.c_int(verbose),
        C.byref(dim), cb.ctypes.data_as(_dp), ce.ctypes.data_as(_dp), cnt.ctypes.data_as(_ip),
        raw.ctypes.data_as(_dp), shot.ctypes.data_as(_dp), C.byref(elapsed)))
    n = dim.value
    raw_c = raw[0:2*n:2] + 1j * raw[1:2*n:2]
    shot_c = shot[0:2*n:2] + 1j * shot[1:2*n:2]
    if stat == "powspec":
        out = {"kbin": cb[:n].copy(), "keff": ce[:n].copy(), "nmodes": cnt[:n].copy(),
               "pk_raw": raw_c, "pk_shot": shot_c}
    else:
        out = {"rbin": cb[:n].copy(), "reff": ce[:n].copy(), "npairs": cnt[:n].copy(),
               "xi": raw_c}
    out["elapsed_s"] = elapsed.value
    return out


def norm_particles_2pt(pos, nz, ws=None, wc=None, alpha=1.):
    """1 / (alpha sum ws nz wc^2), S/twopt.cpp:56-96."""
    return _norm(0, pos, nz, ws, wc, alpha, [1., 1., 1.], [4, 4, 4], "tsc", fn="trvref_norm_powspec")


def norm_mesh_2pt(pos, boxsize, ngrid, assignment, ws=None, wc=None, alpha=1.):
    return _norm(1, pos, None, ws, wc, alpha, boxsize, ngrid, assignment, fn="trvref_norm_powspec")


def bispec_setup(pos, boxsize, ngrid, assignment, bin_range, num_bins):
    """Bin-pair-independent part of the reference's box bispectrum
    (S/threept.cpp:1543-1672); returns its wall time in seconds."""
    pos = np.asarray(pos, dtype=np.float64)
    boxsize = np.broadcast_to(np.asarray(boxsize, dtype=np.float64), (3,)).copy()
    ngrid = np.broadcast_to(np.asarray(ngrid, dtype=np.int32), (3,)).copy()
    x, px = _d(pos[0]); y, py = _d(pos[1]); z, pz = _d(pos[2])
    t = C.c_double(0.)
    _check(lib().trvref_bispec_setup(
        C.c_int(pos.shape[1]), px, py, pz, boxsize.ctypes.data_as(_dp),
        ngrid.ctypes.data_as(_ip), assignment.encode(), C.c_double(bin_range[0]),
        C.c_double(bin_range[1]), C.c_int(num_bins), C.byref(t)))
    return t.value


def bispec_pair(idx_row, idx_col):
    """One bin pair of the reference loop (S/threept.cpp:1900-1968, 2126-2140);
    returns (bk component, shot component, wall seconds)."""
    out = np.zeros(4)
    t = C.c_double(0.)
    _check(lib().trvref_bispec_pair(C.c_int(idx_row), C.c_int(idx_col),
                                    out.ctypes.data_as(_dp), C.byref(t)))
    return complex(out[0], out[1]), complex(out[2], out[3]), t.value


def bispec_pair_shells():
    """(k_eff_a, k_eff_b), (nmodes_a, nmodes_b) of the last :func:`bispec_pair`
    (S/threept.cpp:1912-1919)."""
    keff = np.zeros(2)
    nmodes = np.zeros(2, dtype=np.int32)
    _check(lib().trvref_bispec_pair_shells(keff.ctypes.data_as(_dp), nmodes.ctypes.data_as(_ip)))
    return keff, nmodes


def bispec_twopt(num_bins):
    """Binned ``pk`` and ``sn`` of ``compute_ylm_wgtd_2pt_stats_in_fourier(dn_00,
    N_L0, Sbar, 0, 0)`` (S/threept.cpp:1988-1992); complex arrays of length num_bins."""
    buf = np.zeros(4 * num_bins)
    _check(lib().trvref_bispec_twopt(buf.ctypes.data_as(_dp)))
    buf = buf.reshape(num_bins, 4)
    return buf[:, 0] + 1j * buf[:, 1], buf[:, 2] + 1j * buf[:, 3]


def bispec_entries(pairs, num_bins, ntotal, norm_factor=1.):
    """Entries of the reference's B_000 box data vector for the listed (row, column)
    bin pairs, assembled exactly as compute_bispec_in_gpp_box does for
    (l1, l2, L) = (0, 0, 0) (coupling = 1, no mirror term, unit phase;
    S/threept.cpp:1981-1986, 2044-2056, 2110-2122, 2126-2140, 2162-2163).
    Needs :func:`bispec_setup` first.  Returns a dict shaped like the estimator's
    output restricted to those pairs, plus the wall seconds of every pair unit."""
    pk, sn = bispec_twopt(num_bins)
    sbar = complex(float(ntotal))
    out = {k: [] for k in ("k1_eff", "k2_eff", "nmodes_1", "nmodes_2", "bk_raw", "bk_shot",
                           "unit_s")}
    for a, b in pairs:
        bk, s_ab, t = bispec_pair(a, b)
        keff, nmodes = bispec_pair_shells()
        shot = 0j
        shot += sbar                      # S|{i = j = k}
        shot += pk[a] - sn[a]             # S|{i != j = k} (row bin)
        shot += pk[b] - sn[b]             # S|{j != i = k} (column bin)
        shot += s_ab                      # S|{i = j != k}
        out["k1_eff"].append(keff[0]); out["k2_eff"].append(keff[1])
        out["nmodes_1"].append(nmodes[0]); out["nmodes_2"].append(nmodes[1])
        out["bk_raw"].append(norm_factor * bk); out["bk_shot"].append(norm_factor * shot)
        out["unit_s"].append(t)
    return {k: np.asarray(v) for k, v in out.items()}


def triu_index(a, b, num_bins):
    """Index of pair (a, b), b >= a, in the `triu` data vector (S/threept.cpp:1908-1909)."""
    return (2 * num_bins - a + 1) * a // 2 + (b - a)


def bispec_teardown():
    lib().trvref_bispec_teardown()


def fft_time(n, reps=2):
    """Best-of-``reps`` seconds of one n^3 forward complex transform by the FFTW
    stand-in, through the calls the reference makes (S/field.cpp:246,1552)."""
    t = C.c_double(0.)
    _check(lib().trvref_fft_time(C.c_int(n), C.c_int(reps), C.byref(t)))
    return t.value


def norm_particles(pos, nz, ws=None, wc=None, alpha=1.):
    """``1/(alpha * sum ws nz^2 wc^3)`` (S/threept.cpp:96-136)."""
    return _norm(0, pos, nz, ws, wc, alpha, [1., 1., 1.], [2, 2, 2], "tsc")


def norm_mesh(pos, boxsize, ngrid, assignment, ws=None, wc=None, alpha=1.):
    """Mesh-based normalisation (S/threept.cpp:138-149, S/field.cpp:2017-2065)."""
    return _norm(1, pos, None, ws, wc, alpha, boxsize, ngrid, assignment)


def _norm(from_mesh, pos, nz, ws, wc, alpha, boxsize, ngrid, assignment, fn="trvref_norm"):
    L = lib()
    pos = np.asarray(pos, dtype=np.float64)
    n = pos.shape[1]
    boxsize = np.broadcast_to(np.asarray(boxsize, dtype=np.float64), (3,)).copy()
    ngrid = np.broadcast_to(np.asarray(ngrid, dtype=np.int32), (3,)).copy()
    x, px = _d(pos[0]); y, py = _d(pos[1]); z, pz = _d(pos[2])
    nz_, pnz = _d(nz); ws_, pws = _d(ws); wc_, pwc = _d(wc)
    out = C.c_double(0.)
    _check(getattr(L, fn)(
        C.c_int(from_mesh), C.c_int(n), px, py, pz, pnz, pws, pwc,
        C.c_double(alpha), boxsize.ctypes.data_as(_dp),
        ngrid.ctypes.data_as(_ip), assignment.encode(), C.byref(out)))
    return out.value


def mesh(pos, boxsize, ngrid, assignment, stage=0, subtract_mean=False,
         interlace=False, weights=None, return_time=False):
    """Run the reference MeshField pipeline up to ``stage`` (see ref_driver.cpp)
    and return the complex mesh of shape ``ngrid``."""
    L = lib()
    pos = np.asarray(pos, dtype=np.float64)
    n = pos.shape[1]
    boxsize = np.broadcast_to(np.asarray(boxsize, dtype=np.float64), (3,)).copy()
    ngrid = np.broadcast_to(np.asarray(ngrid, dtype=np.int32), (3,)).copy()
    x, px = _d(pos[0]); y, py = _d(pos[1]); z, pz = _d(pos[2])
    if weights is not None:
        weights = np.asarray(weights, dtype=np.complex128)
        wr, pwr = _d(weights.real); wi, pwi = _d(weights.imag)
    else:
        pwr = pwi = None
    nmesh = int(ngrid[0]) * int(ngrid[1]) * int(ngrid[2])
    out = np.zeros(2 * nmesh)
    t = C.c_double(0.)
    _check(L.trvref_mesh(
        C.c_int(stage), C.c_int(1 if subtract_mean else 0),
        C.c_int(1 if interlace else 0), C.c_int(n), px, py, pz, pwr, pwi,
        boxsize.ctypes.data_as(_dp), ngrid.ctypes.data_as(_ip),
        assignment.encode(), out.ctypes.data_as(_dp), C.byref(t)))
    field = out.view(np.complex128).reshape(tuple(int(v) for v in ngrid))
    return (field, t.value) if return_time else field


def ylm(ell, m, pos):
    """Reduced spherical harmonics (S/maths.cpp:171-220); ``pos`` is (N, 3)."""
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
    out = np.zeros(2 * len(pos))
    lib().trvref_ylm(C.c_int(ell), C.c_int(m), pos.ctypes.data_as(_dp),
                     C.c_int(len(pos)), out.ctypes.data_as(_dp))
    return out.view(np.complex128)


def ylm_mesh(space, ell, m, boxsize, ngrid):
    """``store_reduced_spherical_harmonic_in_{fourier,config}_space`` (S/maths.cpp:222-302)."""
    boxsize = np.broadcast_to(np.asarray(boxsize, dtype=np.float64), (3,)).copy()
    ngrid = np.broadcast_to(np.asarray(ngrid, dtype=np.int32), (3,)).copy()
    out = np.zeros(2 * int(np.prod(ngrid)))
    _check(lib().trvref_ylm_mesh(C.c_int(1 if space == "fourier" else 0), C.c_int(ell), C.c_int(m),
                                 boxsize.ctypes.data_as(_dp), ngrid.ctypes.data_as(_ip),
                                 out.ctypes.data_as(_dp)))
    return out.view(np.complex128).reshape(tuple(int(v) for v in ngrid))


def sjl(ell, x):
    """``SphericalBesselCalculator(ell).eval(x)`` (S/maths.cpp:309-375)."""
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    out = np.zeros(len(x))
    lib().trvref_sjl(C.c_int(ell), x.ctypes.data_as(_dp), C.c_int(len(x)),
                     out.ctypes.data_as(_dp))
    return out


def w3j(j1, j2, j3, m1, m2, m3):
    return lib().trvref_w3j(*(C.c_int(v) for v in (j1, j2, j3, m1, m2, m3)))


def coupling(l1, l2, L_, m1, m2, M):
    return lib().trvref_coupling(*(C.c_int(v) for v in (l1, l2, L_, m1, m2, M)))


def binning(space, scheme, bmin, bmax, nb, boxsize=1000., ngrid=64):
    boxsize = np.broadcast_to(np.asarray(boxsize, dtype=np.float64), (3,)).copy()
    ngrid = np.broadcast_to(np.asarray(ngrid, dtype=np.int32), (3,)).copy()
    edges = np.zeros(nb + 1); centres = np.zeros(nb); widths = np.zeros(nb)
    _check(lib().trvref_binning(
        space.encode(), scheme.encode(), C.c_double(bmin), C.c_double(bmax),
        C.c_int(nb), boxsize.ctypes.data_as(_dp), ngrid.ctypes.data_as(_ip),
        edges.ctypes.data_as(_dp), centres.ctypes.data_as(_dp),
        widths.ctypes.data_as(_dp)))
    return edges, centres, widths


def validate(catalogue_type, statistic_type, assignment="tsc",
             interlace="false", form="diag", degrees=(0, 0, 0), num_bins=4,
             idx_bin=0, bin_range=(0.005, 0.105)):
    bufs = [C.create_string_buffer(64) for _ in range(4)]
    order = C.c_int(0)
    _check(lib().trvref_validate(
        catalogue_type.encode(), statistic_type.encode(), assignment.encode(),
        interlace.encode(), form.encode(), C.c_int(degrees[0]),
        C.c_int(degrees[1]), C.c_int(degrees[2]), C.c_int(num_bins),
        C.c_int(idx_bin), C.c_double(bin_range[0]), C.c_double(bin_range[1]),
        *bufs, C.byref(order)))
    return {
        "shape": bufs[0].value.decode(), "interlace": bufs[1].value.decode(),
        "npoint": bufs[2].value.decode(), "space": bufs[3].value.decode(),
        "assignment_order": order.value,
    }
