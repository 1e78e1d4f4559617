"""The step right before the hot path, on the same nearest-neighbour kernels:

    label_points_by_skeleton(xyz_pcd, xyz_skeleton)   segment_matching/segment_matching_utils.py:128-137 — every point of an
        organ takes the segment id of its nearest skeleton node (KDTree(xyz_skeleton[:, :3]).query(xyz_pcd, k=1), then
        seg_skeleton[nn]); returns (seg_pcd, nn)
    split_by_label(xyz_pcd, seg_pcd, segment_index)    the boolean-mask split of :147-152
    skeleton_landmarks(skel_1, skel_2, voxel_size, sample_method)   segment_matching/segment_matching_ml.py:210-239 — the
        greedy thinning of the two skeleton polylines and the paired landmark sampling that produce lm1 / lm2 of
        get_p2p_mapping_from_segments_using_icp (host code: a few dozen nodes, inherently sequential)
"""
from __future__ import annotations

import numpy as np

from . import _native
from .registration_icp import _points


def label_points_by_skeleton(xyz_pcd, xyz_skeleton, ctx=None):
    """xyz_pcd (N,3); xyz_skeleton (M,4) = x, y, z, segment id  ->  seg_pcd (N,) (dtype of the id column), nn (N,) int64."""
    sk = np.asarray(xyz_skeleton, dtype=np.float64)
    if sk.ndim != 2 or sk.shape[1] < 4:
        raise ValueError(f"xyz_skeleton must have shape (m, 4), got {sk.shape}")
    pts = _points(xyz_pcd, "xyz_pcd")
    nodes = _points(sk[:, :3], "xyz_skeleton[:, :3]")
    ctx = ctx or _native.default_context()
    with ctx.lock:
        _, nn = ctx.nn(pts, nodes)                               # exact 1-NN, lowest index on exact ties
    return sk[:, 3][nn], nn


def split_by_label(xyz_pcd, seg_pcd, segment_index):
    """{seg_i: xyz_pcd[seg_pcd == seg_i]} in the order of segment_index (:147-152)"""
    xyz_pcd = np.asarray(xyz_pcd)
    return {seg_i: xyz_pcd[seg_pcd == seg_i] for seg_i in segment_index}


def _thin(skel, spacing):
    keep = [skel[0, :]]
    for x in skel:                                           # :212-215 — keep a node once it is farther than `spacing` from the last kept
        if np.linalg.norm(keep[-1] - x) > spacing:
            keep.append(x)
    return np.vstack(keep)


def skeleton_landmarks(skel_1, skel_2, voxel_size, sample_method="end"):
    """Landmarks of a matched segment pair from its two skeleton polylines; `voxel_size` is the cleaning voxel size of the
    options file (the reference doubles it, :210).  Returns (lm1, lm2) with the same number of rows."""
    skel_1 = np.asarray(skel_1, dtype=np.float64)[:, :3]
    skel_2 = np.asarray(skel_2, dtype=np.float64)[:, :3]
    spacing = voxel_size * 2
    skel_1, skel_2 = _thin(skel_1, spacing), _thin(skel_2, spacing)
    if sample_method == "end":                               # :225-230
        n = min(skel_1.shape[0], skel_2.shape[0])
        skel_1, skel_2 = skel_1[:n], skel_2[:n]
        return (np.vstack([skel_1[0:skel_1.shape[0] - 1:2], skel_1[-1]]),
                np.vstack([skel_2[0:skel_2.shape[0] - 1:2], skel_2[-1]]))
    len_sample = min(skel_1.shape[0], skel_2.shape[0]) / 2   # :232-239
    step_1, step_2 = int(skel_1.shape[0] / len_sample) + 1, int(skel_2.shape[0] / len_sample) + 1
    skel_1, skel_2 = skel_1[0:skel_1.shape[0]:step_1], skel_2[0:skel_2.shape[0]:step_2]
    n = min(skel_1.shape[0], skel_2.shape[0])
    return skel_1[:n], skel_2[:n]
