"""Drop-in for the reference's ``p2p_matching_in_organ/p2p_matching_evaluation.py`` (the score triple the driver
returns and saves, p2p_matching_seg_icp.py:184-189, run_registration_pipeline.py:143-146):

    evaluate_p2p_matching(mesh1, mesh2, skel_landmarks1, skel_landmarks2, T21, T12=None,
                          voxel_size=1.0, sample_proportion=0.1, verbose=True)      reference :8-88

Same signature, same list of scores [continuity, landmark precision, cycle consistency]; the nearest-neighbour
queries, gathers and counts run on the GPU through liblocalicp.so (a batch of one segment pair).
``evaluate_batch`` scores all segment pairs of a batch in one call.
"""
from __future__ import annotations

import numpy as np

from . import _native
from .registration_icp import _points


def evaluate_batch(batch, T21, T12=None, voxel_size=1.0, sample_proportion=0.1, ctx=None):
    """scores (S,3) float64 = [continuity, landmark precision, cycle] per segment pair of a SegmentPairBatch."""
    ctx = ctx or _native.default_context()
    with ctx.lock:
        ctx.plan(batch.off1, batch.off2, batch.lm_off)
        scores, _ = ctx.evaluate(batch.xyz1, batch.xyz2, batch.lm1, batch.lm2, T21, T12, voxel_size, sample_proportion)
    return scores


def evaluate_p2p_matching(mesh1, mesh2, skel_landmarks1, skel_landmarks2, T21, T12=None, voxel_size=1.0,
                          sample_proportion=0.1, verbose=True):
    v1 = _points(np.asarray(mesh1.vertices), "mesh1.vertices")
    v2 = _points(np.asarray(mesh2.vertices), "mesh2.vertices")
    l1 = np.asarray(skel_landmarks1, np.float64).reshape(-1, 3)
    l2 = np.asarray(skel_landmarks2, np.float64).reshape(-1, 3)
    with_lm = l1.shape[0] > 0 and l2.shape[0] > 0               # reference :23-27
    if with_lm and l1.shape[0] != l2.shape[0]:
        raise ValueError("skel_landmarks1 and skel_landmarks2 must hold the same number of landmarks")
    if not with_lm:                                              # a placeholder landmark; its score is dropped below
        l1, l2 = v1[:1], v2[:1]
    ctx = _native.default_context()
    with ctx.lock:
        ctx.plan([0, len(v1)], [0, len(v2)], [0, len(l1)])
        scores, _ = ctx.evaluate(v1, v2, l1, l2, np.asarray(T21), None if T12 is None else np.asarray(T12),
                                 voxel_size, sample_proportion)
    continuity, lm_precision, cycle = (float(x) for x in scores[0])
    res = [continuity]
    if verbose:
        print("evaluate: continuity -- ", continuity, end=", ")
    if with_lm:
        if verbose:
            print(" landmark precision -- ", lm_precision, end=", ")
        res.append(lm_precision)
    if T12 is not None:
        if verbose:
            print("cycle mesure -- ", cycle)
        res.append(cycle)
    return res
