"""SURVEY section 8d metric (ii): particles assigned per second for one
assignment call (counting sort + zero-fill + scatter), per scheme, with and
without the shifted shadow mesh, throughput and deterministic modes, on the C2
catalogues (1e7 uniform; 1e7 lognormal, P(k) = 2e4 (k/0.05)^-1.5, 256^3 Gaussian
field, default_rng(69)) onto a 512^3 REAL mesh.  Prints one JSON object."""
import ctypes as C
import json
import os
import sys
sys.path.insert(0, '.')
import numpy as np
import torch
from triumvirate_b200 import _lib

dev = torch.device("cuda:0")
tb = _lib.trvb()
tb.trvb_ctx_stream.restype = C.c_void_p
N, L, NG = 10**7, 1000., 512


def lognormal_catalogue(n, L, ngf=256, seed=69):
    gen = np.random.default_rng(seed)
    kf = 2 * np.pi / L
    kx = np.fft.fftfreq(ngf, 1. / ngf) * kf
    kz = np.fft.rfftfreq(ngf, 1. / ngf) * kf
    kk = np.sqrt(kx[:, None, None]**2 + kx[None, :, None]**2 + kz[None, None, :]**2)
    pk = np.zeros_like(kk)
    nz = kk > 0
    pk[nz] = 2.e4 * (kk[nz] / 0.05) ** -1.5
    pk[kk > np.pi * ngf / L] = 0.
    white = np.fft.rfftn(gen.normal(size=(ngf, ngf, ngf)))
    dg = np.fft.irfftn(white * np.sqrt(pk / L**3) * ngf**1.5, s=(ngf, ngf, ngf), axes=(0, 1, 2))
    dln = np.exp(dg - dg.var() / 2.)
    lam = dln * (n / dln.sum())
    cnt = gen.poisson(lam)
    idx = np.repeat(np.arange(ngf**3), cnt.ravel())
    i, j, k = np.unravel_index(idx, (ngf, ngf, ngf))
    cell = L / ngf
    pos = np.stack([(i + gen.uniform(size=i.size)) * cell, (j + gen.uniform(size=i.size)) * cell,
                    (k + gen.uniform(size=i.size)) * cell])
    return np.ascontiguousarray(pos), float(dln.max() / dln.mean())


class Mesh(C.Structure):
    _fields_ = [("data", C.c_void_p), ("layout", C.c_int), ("k0_add", C.c_double)]


def chk(st):
    if st != 0:
        raise RuntimeError(tb.trvb_last_error().decode())


def measure(dpos, order, shifted, mode, reps=5):
    n = dpos.shape[1]
    ctx = C.c_void_p()
    chk(tb.trvb_ctx_create(C.byref(ctx), C.c_int(0), (C.c_int * 3)(NG, NG, NG),
                           (C.c_double * 3)(L, L, L), C.c_int(order)))
    cat = C.c_void_p()
    chk(tb.trvb_cat_create(ctx, C.byref(cat), C.c_longlong(n), C.c_void_p(dpos[0].data_ptr()),
                           C.c_void_p(dpos[1].data_ptr()), C.c_void_p(dpos[2].data_ptr()),
                           None, None, C.c_int(1)))
    mesh = torch.empty(NG**3, dtype=torch.float64, device=dev)
    m = Mesh(mesh.data_ptr(), 0, 0.)
    stream = torch.cuda.ExternalStream(tb.trvb_ctx_stream(ctx), device=dev)

    def call(resort):
        if resort:
            tb.trvb_cat_invalidate_sort(cat)
        chk(tb.trvb_assign(ctx, cat, C.c_int(0), C.c_int(0), C.c_int(0), C.c_double(1.),
                           C.c_int(0), C.c_int(0), C.c_int(shifted), C.c_int(mode), m))

    out = {}
    for key, resort in (("scatter_only_ms", False), ("with_sort_ms", True)):
        for _ in range(2):
            call(resort)
        tb.trvb_ctx_sync(ctx)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            call(resort)
        e1.record(stream)
        tb.trvb_ctx_sync(ctx)
        out[key] = e0.elapsed_time(e1) / reps
    total = float(mesh.sum().item())
    assert abs(total - n) < 1.e-6 * n, (total, n)
    tb.trvb_cat_destroy(cat)
    tb.trvb_ctx_destroy(ctx)
    out["particles_per_s"] = n / (out["with_sort_ms"] * 1.e-3)
    out["alg_GBps_with_sort"] = (32. * n + 8. * NG**3) / (out["with_sort_ms"] * 1.e-3) / 1.e9
    return out


res = {"mesh": NG, "boxsize": L, "unit_note": "REAL 512^3 mesh, unit weights; ms per call"}
uni = np.random.default_rng(42).uniform(0., L, size=(3, N))
logn, contrast = lognormal_catalogue(N, L)
res["lognormal"] = {"n": int(logn.shape[1]), "max_over_mean_density": contrast}
for cname, pos in (("uniform", uni), ("lognormal", logn)):
    d = torch.from_numpy(pos).to(dev)
    for scheme, order in (("ngp", 1), ("cic", 2), ("tsc", 3), ("pcs", 4)):
        for shifted in (0, 1):
            for mode in (0, 1):
                if mode == 1 and (shifted == 1 or cname == "lognormal" and order < 3):
                    continue
                if os.environ.get("SWEEP_ONLY") and os.environ["SWEEP_ONLY"] not in (
                        "deterministic" if mode else "throughput"):
                    continue
                key = f"{cname}/{scheme}/{'shadow' if shifted else 'primary'}/{'deterministic' if mode else 'throughput'}"
                res[key] = measure(d, order, shifted, mode)
                print(key, {k: round(v, 3) if v < 1e6 else f"{v:.3e}" for k, v in res[key].items()},
                      file=sys.stderr, flush=True)
print(json.dumps(res))
