"""Array-level bindings to libtrv_b200.so (src/capi.cpp).

These functions take plain numpy arrays and mirror, argument for argument,
what the reference's Cython layer hands to its C++ estimators
(``T/_threept.pyx:128-248``).  The object-level API (catalogues, parameter
sets) lives in :mod:`triumvirate_b200.threept`.
"""
import ctypes as C

import numpy as np

from ._lib import trv as _trv

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


class TriumvirateError(RuntimeError):
    """A C++ exception crossed the C boundary (status, message)."""


def _check(status):
    if status != 0:
        msg = _trv().trv_last_error().decode()
        if status == 2:
            raise ValueError(msg)
        raise TriumvirateError(msg)


def _d(a):
    if a is None:
        return None, None
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _box(boxsize, ngrid):
    boxsize = np.broadcast_to(np.asarray(boxsize, dtype=np.float64), (3,)).copy()
    ngrid = np.broadcast_to(np.asarray(ngrid, dtype=np.int32), (3,)).copy()
    return boxsize, ngrid


def gpu_count():
    return _trv().trv_gpu_count()


def partition_owners(form, degrees, num_bins, world, idx_bin=0):
    """Owner rank of every data-vector entry (trv::partition_owners)."""
    owner = np.zeros(max(num_bins * num_bins, num_bins), dtype=np.int32)
    dim = C.c_int(0)
    st = _trv().trv_partition_owners(form.encode(), C.c_int(degrees[0]), C.c_int(degrees[1]),
                                     C.c_int(idx_bin), C.c_int(num_bins), C.c_int(world),
                                     owner.ctypes.data_as(C.c_void_p), C.byref(dim))
    _check(st)
    return owner[:dim.value].copy()


def counters():
    f = C.c_int(0); i = C.c_int(0); g = C.c_double(0.)
    _trv().trv_counters(C.byref(f), C.byref(i), C.byref(g))
    return {"count_fft": f.value, "count_ifft": i.value, "gib_gpu_max": g.value}


def twopt(stat, catalogue_type, boxsize, ngrid, assignment, degree, bin_range, num_bins,
          norm_factor, pos_d=None, nz_d=None, ws_d=None, wc_d=None, los_d=None,
          pos_r=None, nz_r=None, ws_r=None, wc_r=None, los_r=None,
          interlace="false", binning="lin", custom_edges=None, alpha=1., verbose=60,
          deterministic=False):
    """Run a two-point estimator on the GPU.

    ``stat`` is ``'powspec'``, ``'2pcf'`` or ``'2pcf-win'``; ``catalogue_type`` is
    ``'sim'`` (``trv::compute_*_in_gpp_box``), ``'survey'`` (``trv::compute_powspec``
    / ``trv::compute_corrfunc``) or ``'random'`` (``trv::compute_corrfunc_window``:
    the catalogue is ``pos_r`` with ``alpha``).  Returns a dict with the
    reference's result-struct field names.
    """
    L = _trv()
    boxsize, ngrid = _box(boxsize, ngrid)
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
    if isinstance(interlace, bool):
        interlace = "true" if interlace else "false"
    status = L.trv_twopt(
        stat.encode(), catalogue_type.encode(),
        C.c_int(nd), *args_d, C.c_int(nr), *args_r,
        boxsize.ctypes.data_as(_dp), ngrid.ctypes.data_as(_ip),
        assignment.encode(), interlace.encode(), C.c_int(degree), binning.encode(),
        C.c_double(bin_range[0]), C.c_double(bin_range[1]), C.c_int(nb), arr(custom_edges),
        C.c_double(alpha), C.c_double(norm_factor), C.c_int(verbose),
        C.c_int(1 if deterministic else 0),
        C.byref(dim), cb.ctypes.data_as(_dp), ce.ctypes.data_as(_dp), cnt.ctypes.data_as(_ip),
        raw.ctypes.data_as(_dp), shot.ctypes.data_as(_dp), C.byref(elapsed),
    )
    _check(status)
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


def threept(stat, catalogue_type, pos_d, boxsize, ngrid, assignment, degrees,
            form, bin_range, num_bins, norm_factor, idx_bin=0, binning="lin",
            nz_d=None, ws_d=None, wc_d=None, los_d=None,
            pos_r=None, nz_r=None, ws_r=None, wc_r=None, los_r=None,
            interlace="false", custom_edges=None, verbose=60,
            deterministic=False, part_rank=0, part_count=1):
    """Run a three-point estimator on the GPU.

    ``stat`` is ``'bispec'`` or ``'3pcf'``; ``catalogue_type`` is ``'sim'``
    (``trv::compute_*_in_gpp_box``) or ``'survey'`` (``trv::compute_bispec`` /
    ``trv::compute_3pcf``).  Positions are ``(3, N)`` arrays already aligned
    in the box; LOS arrays are ``(N, 3)``.  Returns a dict of numpy arrays with
    the reference's result-struct field names.
    """
    L = _trv()
    boxsize, ngrid = _box(boxsize, ngrid)
    keep = []

    def arr(a):
        a_, p = _d(a)
        keep.append(a_)
        return p

    pos_d = np.asarray(pos_d, dtype=np.float64)
    nd = pos_d.shape[1]
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
    if isinstance(interlace, bool):
        interlace = "true" if interlace else "false"
    status = L.trv_threept(
        stat.encode(), catalogue_type.encode(),
        C.c_int(nd), *args_d, C.c_int(nr), *args_r,
        boxsize.ctypes.data_as(_dp), ngrid.ctypes.data_as(_ip),
        assignment.encode(), interlace.encode(),
        C.c_int(degrees[0]), C.c_int(degrees[1]), C.c_int(degrees[2]),
        form.encode(), C.c_int(idx_bin or 0), binning.encode(),
        C.c_double(bin_range[0]), C.c_double(bin_range[1]), C.c_int(nb),
        arr(custom_edges),
        C.c_double(norm_factor), C.c_int(verbose),
        C.c_int(1 if deterministic else 0), C.c_int(part_rank), C.c_int(part_count),
        C.byref(dim),
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
                   wa_orders=(0, 0), verbose=60,
                   custom_edges=None, deterministic=False, part_rank=0, part_count=1):
    """3PCF window function of a random catalogue on the GPU
    (``trv::compute_3pcf_window``, binding ``T/_threept.pyx:276-314``).  ``pos_r`` is
    ``(3, N)`` already aligned in the box, ``los_r`` is ``(N, 3)``."""
    L = _trv()
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
    status = L.trv_threept_window(
        C.c_int(pos_r.shape[1]), *ptrs,
        boxsize.ctypes.data_as(_dp), ngrid.ctypes.data_as(_ip), assignment.encode(),
        C.c_int(degrees[0]), C.c_int(degrees[1]), C.c_int(degrees[2]),
        C.c_int(wa_orders[0]), C.c_int(wa_orders[1]),
        form.encode(), C.c_int(idx_bin or 0), binning.encode(),
        C.c_double(bin_range[0]), C.c_double(bin_range[1]), C.c_int(nb),
        (np.ascontiguousarray(custom_edges, dtype=np.float64).ctypes.data_as(_dp)
         if custom_edges is not None else None),
        C.c_double(alpha), C.c_double(norm_factor), C.c_int(1 if wide_angle else 0),
        C.c_int(verbose),
        C.c_int(1 if deterministic else 0), C.c_int(part_rank), C.c_int(part_count),
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


def threept_box_arrays(stat, n, x_ptr, y_ptr, z_ptr, on_device, boxsize, ngrid,
                       assignment, degrees, form, bin_range, num_bins,
                       norm_factor=1., idx_bin=0, binning="lin", verbose=60,
                       deterministic=False, part_rank=0, part_count=1):
    """Periodic-box estimator from three coordinate arrays given by ADDRESS
    (``x_ptr`` etc. are integers: host addresses, e.g. of pinned numpy/torch
    buffers, or CUDA device addresses when ``on_device``).  Unit weights."""
    L = _trv()
    boxsize, ngrid = _box(boxsize, ngrid)
    nb = int(num_bins)
    cap = max(nb * nb, nb) + 8
    dim = C.c_int(0)
    c1b = np.zeros(cap); c2b = np.zeros(cap)
    c1e = np.zeros(cap); c2e = np.zeros(cap)
    n1 = np.zeros(cap, dtype=np.int32); n2 = np.zeros(cap, dtype=np.int32)
    raw = np.zeros(2 * cap); shot = np.zeros(2 * cap)
    status = L.trv_threept_box_arrays(
        stat.encode(), C.c_longlong(n), C.c_void_p(x_ptr), C.c_void_p(y_ptr),
        C.c_void_p(z_ptr), C.c_int(1 if on_device else 0),
        boxsize.ctypes.data_as(_dp), ngrid.ctypes.data_as(_ip), assignment.encode(),
        C.c_int(degrees[0]), C.c_int(degrees[1]), C.c_int(degrees[2]),
        form.encode(), C.c_int(idx_bin or 0), binning.encode(),
        C.c_double(bin_range[0]), C.c_double(bin_range[1]), C.c_int(nb),
        C.c_double(norm_factor), C.c_int(verbose),
        C.c_int(1 if deterministic else 0), C.c_int(part_rank), C.c_int(part_count),
        C.byref(dim),
        c1b.ctypes.data_as(_dp), c2b.ctypes.data_as(_dp),
        c1e.ctypes.data_as(_dp), c2e.ctypes.data_as(_dp),
        n1.ctypes.data_as(_ip), n2.ctypes.data_as(_ip),
        raw.ctypes.data_as(_dp), shot.ctypes.data_as(_dp),
    )
    _check(status)
    n_ = dim.value
    raw_c = raw[0:2*n_:2] + 1j * raw[1:2*n_:2]
    shot_c = shot[0:2*n_:2] + 1j * shot[1:2*n_:2]
    if stat == "bispec":
        names = ("k1_bin", "k2_bin", "k1_eff", "k2_eff", "nmodes_1",
                 "nmodes_2", "bk_raw", "bk_shot")
    else:
        names = ("r1_bin", "r2_bin", "r1_eff", "r2_eff", "npairs_1",
                 "npairs_2", "zeta_raw", "zeta_shot")
    vals = (c1b[:n_].copy(), c2b[:n_].copy(), c1e[:n_].copy(), c2e[:n_].copy(),
            n1[:n_].copy(), n2[:n_].copy(), raw_c, shot_c)
    return dict(zip(names, vals))


def _prefer_bundled_nccl():
    """Point TRV_NCCL_LIB at the libnccl.so.2 that ships with PyTorch (pip package
    nvidia-nccl), so that this process maps one NCCL only -- the one torch itself needs --
    whichever of the two libraries asks for it first."""
    import os
    if os.environ.get("TRV_NCCL_LIB"):
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for root in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(root, "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["TRV_NCCL_LIB"] = cand
                return
    except Exception:
        pass


def comm_unique_id():
    """128-byte NCCL id created by rank 0 (``trv_comm_unique_id``)."""
    _prefer_bundled_nccl()
    buf = C.create_string_buffer(128)
    _check(_trv().trv_comm_unique_id(buf))
    return buf.raw


def comm_init(nranks, rank, ident):
    """Attach an NCCL communicator to this process (collective, ``trv_comm_init``)."""
    _prefer_bundled_nccl()
    _check(_trv().trv_comm_init(C.c_int(nranks), C.c_int(rank), C.c_char_p(bytes(ident))))


def comm_size():
    return _trv().trv_comm_size()


def comm_finalize():
    _trv().trv_comm_finalize()


def allreduce(buf):
    """In-place sum of a float64 array over the ranks of the process communicator."""
    buf = np.ascontiguousarray(buf, dtype=np.float64)
    _check(_trv().trv_allreduce(buf.ctypes.data_as(_dp), C.c_longlong(buf.size)))
    return buf


def multi_device_count(ngrid):
    """GPUs a single-process estimator call on this mesh spreads over."""
    ngrid = np.broadcast_to(np.asarray(ngrid, dtype=np.int32), (3,)).copy()
    return _trv().trv_multi_device_count(ngrid.ctypes.data_as(_ip))


def dmesh_call_count():
    """Estimator calls of this process that ran the distributed mesh phase."""
    fn = _trv().trv_dmesh_call_count
    fn.restype = C.c_longlong
    return int(fn())


def fused_mesh_call_count():
    """Estimator calls of this process whose mesh phase ran with the fused x pass."""
    fn = _trv().trv_fused_mesh_call_count
    fn.restype = C.c_longlong
    return int(fn())


def release_contexts():
    """Drop the cached device contexts (cuFFT plans, tables) and hand the arena's
    cached blocks back to the driver."""
    _trv().trv_release_contexts()


def profile_enable(on=True):
    _trv().trv_profile_enable(C.c_int(1 if on else 0))


def profile_report():
    import json
    buf = C.create_string_buffer(4096)
    _trv().trv_profile_report(buf, C.c_int(4096))
    return json.loads(buf.value.decode() or "{}")


def norm_particles(pos, nz, ws=None, wc=None, alpha=1.):
    return _norm(0, pos, nz, ws, wc, alpha, [1., 1., 1.], [4, 4, 4], "tsc")


def norm_mesh(pos, boxsize, ngrid, assignment, ws=None, wc=None, alpha=1.):
    return _norm(1, pos, None, ws, wc, alpha, boxsize, ngrid, assignment)


def norm_particles_2pt(pos, nz, ws=None, wc=None, alpha=1.):
    return _norm(0, pos, nz, ws, wc, alpha, [1., 1., 1.], [4, 4, 4], "tsc", fn="trv_norm_powspec")


def norm_mesh_2pt(pos, boxsize, ngrid, assignment, ws=None, wc=None, alpha=1.):
    return _norm(1, pos, None, ws, wc, alpha, boxsize, ngrid, assignment, fn="trv_norm_powspec")


def _norm(from_mesh, pos, nz, ws, wc, alpha, boxsize, ngrid, assignment, fn="trv_norm"):
    pos = np.asarray(pos, dtype=np.float64)
    n = pos.shape[1]
    boxsize, ngrid = _box(boxsize, ngrid)
    x, px = _d(pos[0]); y, py = _d(pos[1]); z, pz = _d(pos[2])
    nz_, pnz = _d(nz); ws_, pws = _d(ws); wc_, pwc = _d(wc)
    out = C.c_double(0.)
    _check(getattr(_trv(), fn)(
        C.c_int(from_mesh), C.c_int(n), px, py, pz, pnz, pws, pwc,
        C.c_double(alpha), boxsize.ctypes.data_as(_dp),
        ngrid.ctypes.data_as(_ip), assignment.encode(), C.byref(out)))
    return out.value


def mesh(pos, boxsize, ngrid, assignment, stage=0, subtract_mean=False,
         interlace=False, weights=None, deterministic=False, return_time=False):
    """``trv::MeshField`` pipeline up to ``stage`` (see ``trv_mesh``); returns
    the complex mesh with shape ``ngrid``."""
    pos = np.asarray(pos, dtype=np.float64)
    n = pos.shape[1]
    boxsize, ngrid = _box(boxsize, ngrid)
    x, px = _d(pos[0]); y, py = _d(pos[1]); z, pz = _d(pos[2])
    if weights is not None:
        weights = np.asarray(weights, dtype=np.complex128)
        wr, pwr = _d(weights.real); wi, pwi = _d(weights.imag)
    else:
        pwr = pwi = None
    nmesh = int(ngrid[0]) * int(ngrid[1]) * int(ngrid[2])
    out = np.zeros(2 * nmesh)
    t = C.c_double(0.)
    _check(_trv().trv_mesh(
        C.c_int(stage), C.c_int(1 if subtract_mean else 0),
        C.c_int(1 if interlace else 0), C.c_int(1 if deterministic else 0),
        C.c_int(n), px, py, pz, pwr, pwi,
        boxsize.ctypes.data_as(_dp), ngrid.ctypes.data_as(_ip),
        assignment.encode(), out.ctypes.data_as(_dp), C.byref(t)))
    field = out.view(np.complex128).reshape(tuple(int(v) for v in ngrid))
    return (field, t.value) if return_time else field


def ylm(ell, m, pos):
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
    out = np.zeros(2 * len(pos))
    _trv().trv_ylm(C.c_int(ell), C.c_int(m), pos.ctypes.data_as(_dp),
                   C.c_int(len(pos)), out.ctypes.data_as(_dp))
    return out.view(np.complex128)


def ylm_mesh(space, ell, m, boxsize, ngrid):
    """``store_reduced_spherical_harmonic_in_{fourier,config}_space`` tables (host)."""
    boxsize, ngrid = _box(boxsize, ngrid)
    out = np.zeros(2 * int(np.prod(ngrid)))
    _check(_trv().trv_ylm_mesh(C.c_int(1 if space == "fourier" else 0), C.c_int(ell), C.c_int(m),
                               boxsize.ctypes.data_as(_dp), ngrid.ctypes.data_as(_ip),
                               out.ctypes.data_as(_dp)))
    return out.view(np.complex128).reshape(tuple(int(v) for v in ngrid))


def sjl(ell, x):
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    out = np.zeros(len(x))
    _trv().trv_sjl(C.c_int(ell), x.ctypes.data_as(_dp), C.c_int(len(x)),
                   out.ctypes.data_as(_dp))
    return out


def sjl_exact(ell, x):
    return _trv().trv_sjl_exact(C.c_int(ell), C.c_double(x))


def w3j(j1, j2, j3, m1, m2, m3):
    return _trv().trv_w3j(*(C.c_int(v) for v in (j1, j2, j3, m1, m2, m3)))


def coupling(l1, l2, L_, m1, m2, M):
    return _trv().trv_coupling(*(C.c_int(v) for v in (l1, l2, L_, m1, m2, M)))


def binning(space, scheme, bmin, bmax, nb, boxsize=1000., ngrid=64):
    boxsize, ngrid = _box(boxsize, ngrid)
    edges = np.zeros(nb + 1); centres = np.zeros(nb); widths = np.zeros(nb)
    _check(_trv().trv_binning(
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
    _check(_trv().trv_validate(
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
