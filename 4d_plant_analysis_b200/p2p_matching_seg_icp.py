"""Drop-in for the two hot functions of the reference's
``p2p_matching_in_organ/p2p_matching_seg_icp.py``:

    get_triangle_angles(a, b, c)                                          reference :27-34
    get_p2p_mapping_from_segments_using_icp(mesh1, mesh2, lm1, lm2, vs)   reference :383-489

Same signatures and return types; the work is one batched GPU pass (batch of one segment pair,
one direction) through liblocalicp.so.
"""
from __future__ import annotations

import numpy as np

from . import _native
from .registration_icp import _points


def get_triangle_angles(a, b, c):
    assert a + b >= c and a + c >= b and b + c >= a
    ctx = _native.default_context()
    with ctx.lock:
        return ctx.triangle_angles(np.array([[a, b, c]], dtype=np.float64))[0]


def _vector3d(colors):
    try:                                    # the reference wraps colours in open3d's Vector3dVector
        import open3d
        return open3d.utility.Vector3dVector(colors)
    except Exception:
        return colors


def get_p2p_mapping_from_segments_using_icp(mesh1, mesh2, lm1, lm2, voxel_size):
    """T21 (N2,) int64: for every vertex of mesh2 the index of its match among the vertices of mesh1.
    ``voxel_size`` is accepted and ignored, as in the reference."""
    # (the reference deep-copies both meshes, :384-385, only to protect its inputs; nothing here writes to the vertices, and
    # _points returns a contiguous float64 array — a copy whenever the mesh does not already hold one)
    xyz1 = _points(np.asarray(mesh1.vertices), "mesh1.vertices")
    xyz2 = _points(np.asarray(mesh2.vertices), "mesh2.vertices")
    lm1 = _points(lm1, "lm1")
    lm2 = _points(lm2, "lm2")
    if lm1.shape[0] != lm2.shape[0]:
        raise ValueError("lm1 and lm2 must hold the same number of landmarks")
    ctx = _native.default_context()
    with ctx.lock:                          # plan + run + getters of the shared context belong together
        ctx.plan([0, len(xyz1)], [0, len(xyz2)], [0, len(lm1)], _native.DIR_T21)
        T21, _ = ctx.run_host(xyz1, xyz2, lm1, lm2)
        if lm1.shape[0] > 1:
            # side effects of the reference (:394-398): one draw from the global numpy RNG, vertex colours
            # by nearest landmark
            color_setting = np.random.random((lm1.shape[0], 3))
            try:
                nn1, _ = ctx.landmark_nn(1, distances=False)
                nn2, _ = ctx.landmark_nn(2, distances=False)
                mesh1.vertex_colors = _vector3d(color_setting[nn1[:, 0]])
                mesh2.vertex_colors = _vector3d(color_setting[nn2[:, 0]])
            except AttributeError:
                pass
    return T21
