"""Drop-in for the reference's ``registration/registration_icp.py``: same names, argument meaning,
return shapes and dtypes; the arithmetic runs in liblocalicp.so on the GPU (no CPU path).

    get_best_fit_transform(A, B) -> (T, R, t)      reference registration_icp.py:8-46
    get_NN_in_target(xyz1, xyz2) -> (dist, idx)    reference registration_icp.py:49-57
    my_icp(xyz1, xyz2)           -> (T, R, t)      reference registration_icp.py:60-98
"""
from __future__ import annotations

import numpy as np

from . import _native


def _points(a, name):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError(f"{name} must have shape (n, 3), got {a.shape}")
    if a.shape[0] == 0:
        raise ValueError(f"Found array with 0 sample(s) (shape={a.shape}) while a minimum of 1 is required.")
    if not np.isfinite(a).all():
        raise ValueError("Input contains NaN or infinity.")     # what sklearn's KDTree raises in the reference
    return np.ascontiguousarray(a)


def get_best_fit_transform(A, B):
    """Least-squares rigid fit A -> B (Kabsch).  Returns T (4x4 with R^T top-left and t in the last
    row, as the reference builds it), R (3,3), t (3,)."""
    assert len(A) == len(B)
    A, B = _points(A, "A"), _points(B, "B")
    ctx = _native.default_context()
    with ctx.lock:
        R, t, _ = ctx.kabsch(A, B)
    T = np.identity(4)
    T[0:3, 0:3] = R.T
    T[[3], 0:3] = t.T
    return T, R, t


def get_NN_in_target(xyz1, xyz2):
    """Exact nearest neighbour in xyz2 of every point of xyz1: (distance (n,1) float64, indices (n,) int64)."""
    xyz1, xyz2 = _points(xyz1, "xyz1"), _points(xyz2, "xyz2")
    ctx = _native.default_context()
    with ctx.lock:
        d, i = ctx.nn(xyz1, xyz2)
    return d.reshape(-1, 1), i


def my_icp(xyz1, xyz2):
    """Point-to-point ICP of source xyz1 onto target xyz2.  Returns T (4x4, R top-left, t^T in row 3),
    R (3,3), t (3,1); callers apply it as ``x @ R.T + t.T``."""
    xyz1, xyz2 = _points(xyz1, "xyz1"), _points(xyz2, "xyz2")
    ctx = _native.default_context()
    with ctx.lock:
        R, t, _, _ = ctx.icp(xyz1, [0, len(xyz1)], xyz2, [0, len(xyz2)])
    R, t = R[0], t[0].reshape(3, 1)
    T = np.identity(4)
    T[0:3, 0:3] = R
    T[[3], 0:3] = t.T
    return T, R, t
