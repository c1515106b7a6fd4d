"""Catalogue pre-processing with the reference's Python-side semantics
(``T/catalogue.py``): lines of sight, centring, padding, periodising.

Functions operate on ``(3, N)`` position arrays; :class:`ParticleCatalogue`
wraps them in the reference's object interface.
"""
import warnings

import numpy as np


def compute_los(pos):
    """``pos / |pos|`` with a zero-norm guard (T/catalogue.py:437-458); (N, 3)."""
    pos = np.asarray(pos, dtype=np.float64)
    norm = np.sqrt(pos[0]**2 + pos[1]**2 + pos[2]**2)
    norm[norm == 0.] = 1.
    return np.ascontiguousarray(
        np.transpose([pos[0] / norm, pos[1] / norm, pos[2] / norm])
    )


def _box(boxsize):
    return np.broadcast_to(np.asarray(boxsize, dtype=np.float64), (3,))


def periodise(pos, boxsize):
    """``(x + L/2 - (max + min)/2) % L`` per axis (T/catalogue.py:648-676):
    note that this also centres the catalogue in the box."""
    boxsize = _box(boxsize)
    pos = np.array(pos, dtype=np.float64, copy=True)
    for ax in range(3):
        lo, hi = pos[ax].min(), pos[ax].max()
        pos[ax] = (pos[ax] + boxsize[ax] / 2. - (hi + lo) / 2.) % boxsize[ax]
    return pos


def centre(pos, pos_ref=None, boxsize=None):
    """Shift so that the reference catalogue's extent mid-point sits at the
    box centre (T/catalogue.py:490-545).  Returns the shifted array(s)."""
    boxsize = _box(boxsize)
    pos = np.array(pos, dtype=np.float64, copy=True)
    ref = pos if pos_ref is None else np.array(pos_ref, dtype=np.float64, copy=True)
    origin = np.array([
        np.mean([ref[ax].min(), ref[ax].max()]) - boxsize[ax] / 2. for ax in range(3)
    ])
    pos = pos - origin[:, None] if pos_ref is not None else None
    ref = ref - origin[:, None]
    return ref if pos_ref is None else (pos, ref)


def pad(pos, pos_ref=None, boxsize=None, ngrid=None, boxsize_pad=None, ngrid_pad=None):
    """Shift so that the reference catalogue's minimum corner sits at the
    requested padding from the origin (T/catalogue.py:547-646)."""
    if boxsize_pad is None and ngrid_pad is None:
        warnings.warn("`boxsize_pad` and `ngrid_pad` are both None. No padding is applied.")
        return pos if pos_ref is None else (pos, pos_ref)
    if boxsize_pad is not None and ngrid_pad is not None:
        raise ValueError(
            "Conflicting padding as `boxsize_pad` and `ngrid_pad` are both set (not None).")
    boxsize = _box(boxsize)
    pos = np.array(pos, dtype=np.float64, copy=True)
    ref = pos if pos_ref is None else np.array(pos_ref, dtype=np.float64, copy=True)
    origin = np.array([ref[ax].min() for ax in range(3)])
    if boxsize_pad:
        origin -= np.multiply(boxsize_pad, boxsize)
    if ngrid_pad:
        origin -= np.multiply(ngrid_pad, np.divide(boxsize, ngrid))
    if pos_ref is None:
        return ref - origin[:, None]
    return pos - origin[:, None], ref - origin[:, None]
