"""Imports the UNMODIFIED reference hot-path modules from /root/reference under sys.modules stubs
(SURVEY.md App. B). Only usable in the authoring container — the GPU box has no /root/reference —
so nothing under tests/ imports this at test time; it is used by make_golden.py, which writes the
committed fixtures in this directory.
"""
import os
import sys
import types
from unittest.mock import MagicMock

REF_ROOT = os.environ.get("REF_ROOT", "/root/reference")


class Mesh:
    """Stand-in for open3d.geometry.TriangleMesh: the hot path reads .vertices, writes .vertex_colors."""

    def __init__(self, v):
        self.vertices, self.vertex_colors = v, None


def load():
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"{REF_ROOT} not present (reference loader is authoring-container only)")
    o3d = types.ModuleType("open3d")
    _C = type("_C", (), {})
    o3d.geometry = types.SimpleNamespace(PointCloud=_C, LineSet=_C, TriangleMesh=_C)
    o3d.utility = types.SimpleNamespace(Vector3dVector=lambda a: a, Vector2iVector=lambda a: a)
    sys.modules.setdefault("open3d", o3d)
    for m in ["matplotlib", "matplotlib.pyplot", "dijkstar", "pyntcloud", "spharapy", "spharapy.trimesh",
              "spharapy.spharabasis", "polyscope", "robust_laplacian", "imageio"]:
        sys.modules.setdefault(m, MagicMock(name=m))
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    cwd = os.getcwd()
    os.chdir(os.path.join(REF_ROOT, "p2p_matching_in_organ"))
    try:
        import registration.registration_icp as RI
        import p2p_matching_in_organ.p2p_matching_seg_icp as SI
        import p2p_matching_in_organ.p2p_matching_fm as FM
        import p2p_matching_in_organ.p2p_matching_evaluation as EV
    finally:
        os.chdir(cwd)
    return types.SimpleNamespace(RI=RI, SI=SI, FM=FM, EV=EV, Mesh=Mesh)
