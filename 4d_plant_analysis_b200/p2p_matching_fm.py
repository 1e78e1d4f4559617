"""Drop-in for ``double_direction_refinement`` of the reference's ``p2p_matching_in_organ/p2p_matching_fm.py:75-132``
(called right after the hot path, p2p_matching_seg_icp.py:181-182): same signature, T21 refined in place and
returned together with the unchanged T12.  ``refine_batch`` does all segment pairs of a batch in one call."""
from __future__ import annotations

import numpy as np

from . import _native
from .registration_icp import _points


def refine_batch(batch, T21, T12, ctx=None):
    """refined copy of the concatenated T21 of a SegmentPairBatch (T12 is only read)."""
    ctx = ctx or _native.default_context()
    if (np.diff(batch.off1) < 5).any():
        raise ValueError("Expected n_neighbors <= n_samples_fit: every mesh1 needs at least 5 vertices (KDTree.query(k=5))")
    with ctx.lock:
        ctx.plan(batch.off1, batch.off2, batch.lm_off)
        return ctx.refine(batch.xyz1, batch.xyz2, T21, T12)


def double_direction_refinement(mesh1, mesh2, T21, T12, lm1=None, lm2=None):
    v1 = _points(np.asarray(mesh1.vertices), "mesh1.vertices")
    v2 = _points(np.asarray(mesh2.vertices), "mesh2.vertices")
    if len(v1) < 5:
        raise ValueError(f"Expected n_neighbors <= n_samples_fit, but n_neighbors = 5, n_samples_fit = {len(v1)}")
    ctx = _native.default_context()
    with ctx.lock:
        ctx.plan([0, len(v1)], [0, len(v2)], [0, 1])
        out = ctx.refine(v1, v2, T21, T12)
    if isinstance(T21, np.ndarray) and T21.shape == out.shape:
        T21[...] = out                                       # the reference rewrites the caller's array in place
        return T21, T12
    return out, T12
