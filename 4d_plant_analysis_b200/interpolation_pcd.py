"""Drop-in for ``interpolation_segment_level`` of the reference's ``interpolation/interpolation_pcd.py:122-134`` — the
consumer of the saved maps (``interpolation_plant_level`` :71-76 calls it once per matched segment pair):

    interpolation_segment_level(xyz1, xyz2, T21, T12, alpha) -> (n2 + n1, 3) float64
        rows 0..n2-1   (1 - alpha) * xyz1[T21] + alpha * xyz2
        rows n2..      (1 - alpha) * xyz1      + alpha * xyz2[T12]

Same signature and the same rows bit for bit (the products and the sum are rounded separately on the GPU as numpy
does).  ``T21`` / ``T12`` may be the float columns ``np.loadtxt`` returns for ``T21_{i}.csv`` (the reference casts
with ``.astype(int)``).  ``interpolate_batch`` does every segment pair of a batch in one call;
``interpolate_saved_result`` reads the directory the driver (or ``save_segment_results``) wrote and returns the stacked
rows of all matched segment pairs (the ``np.vstack(P_inter_collection)`` of :76-86).
"""
from __future__ import annotations

import os

import numpy as np

from . import _native


def _points(a, name):                                        # no finiteness check: the reference interpolates whatever it is given
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError(f"{name} must have shape (n, 3), got {a.shape}")
    return np.ascontiguousarray(a)


def _maps(T, name):
    T = np.asarray(T)
    if T.ndim != 1:
        T = T.reshape(-1)
    if T.dtype.kind == "f":
        if not np.isfinite(T).all():
            raise ValueError(f"{name} contains NaN or infinity")
        T = T.astype(int)                                    # reference :128, :131
    return np.ascontiguousarray(T, np.int64)


def interpolate_batch(batch, T21, T12, alpha, ctx=None):
    """(off1[S] + off2[S], 3) float64: per segment pair s, from row off1[s] + off2[s], its interpolation_segment_level rows."""
    ctx = ctx or _native.default_context()
    with ctx.lock:
        ctx.plan(batch.off1, batch.off2, batch.lm_off)
        return ctx.interpolate(batch.xyz1, batch.xyz2, _maps(T21, "T21"), _maps(T12, "T12"), alpha)


def interpolation_segment_level(xyz1, xyz2, T21, T12, alpha):
    v1, v2 = _points(xyz1, "xyz1"), _points(xyz2, "xyz2")
    T21, T12 = _maps(T21, "T21"), _maps(T12, "T12")
    if len(T21) != len(v2) or len(T12) != len(v1):
        raise ValueError("operands could not be broadcast together: T21 must have len(xyz2) entries, T12 len(xyz1)")
    if len(v1) == 0 and len(v2) == 0:
        return np.empty((0, 3))
    if len(v1) == 0 or len(v2) == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    ctx = _native.default_context()
    try:
        with ctx.lock:
            ctx.plan([0, len(v1)], [0, len(v2)], [0, 1])
            return ctx.interpolate(v1, v2, T21, T12, alpha)
    except ValueError as e:                                  # numpy's exception type for a bad fancy index
        raise IndexError(str(e)) from None


def interpolate_saved_result(directory, day1, day2, alpha, ctx=None):
    """All matched segment pairs of one saved registration result (files of p2p_matching_seg_icp.py:210-242, read as
    load_pcd_registration does, interpolation_pcd.py:15-41) interpolated in one GPU call."""
    x1, x2, t21, t12 = [], [], [], []
    i = 0
    while i < 500:                                           # the reference's cap (:25)
        f1 = os.path.join(directory, f"{day1}_{i}.csv")
        if not os.path.exists(f1):
            break
        x1.append(np.loadtxt(f1).reshape(-1, 3))
        x2.append(np.loadtxt(os.path.join(directory, f"{day2}_{i}.csv")).reshape(-1, 3))
        t21.append(np.loadtxt(os.path.join(directory, f"T21_{i}.csv")).reshape(-1))
        t12.append(np.loadtxt(os.path.join(directory, f"T12_{i}.csv")).reshape(-1))
        i += 1
    if not x1:
        return np.empty((0, 3))
    off1 = np.concatenate([[0], np.cumsum([len(a) for a in x1])]).astype(np.int64)
    off2 = np.concatenate([[0], np.cumsum([len(a) for a in x2])]).astype(np.int64)
    ctx = ctx or _native.default_context()
    with ctx.lock:
        ctx.plan(off1, off2, np.arange(len(x1) + 1, dtype=np.int64))
        return ctx.interpolate(np.vstack(x1), np.vstack(x2), _maps(np.concatenate(t21), "T21"), _maps(np.concatenate(t12), "T12"), alpha)
