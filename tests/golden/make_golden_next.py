"""Writes tests/golden/next.npz: outputs of the UNMODIFIED reference for the rows of SURVEY.md §8f that sit either side of
the hot path (f3 interpolation_segment_level, f4 skeleton labelling and landmark sampling).  Authoring container only.

    python tests/golden/make_golden_next.py

* interpolation_segment_level is imported and called as is (interpolation/interpolation_pcd.py:122-134) on segment pairs
  and refined maps of c1_small.npz, once with the maps as the float columns np.loadtxt gives the reference.
* The labelling (segment_matching/segment_matching_utils.py:130-137) and the landmark sampling
  (segment_matching/segment_matching_ml.py:210-239) are statements inside long functions that need Open3D and the ROMI
  data; the very source lines are cut out of the reference files and exec'd on synthetic inputs — the reference's own
  code decides the recorded values, nothing is restated here.
"""
import os
import sys
import textwrap
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import reference_loader  # noqa: E402

ref = reference_loader.load()
REF = reference_loader.REF_ROOT
cwd = os.getcwd()
os.chdir(os.path.join(REF, "interpolation"))
try:
    import interpolation.interpolation_pcd as IP
finally:
    os.chdir(cwd)
from sklearn.neighbors import KDTree  # noqa: E402


def ref_lines(path, first, last):
    """source lines first..last (1-based, inclusive) of a reference file, dedented"""
    with open(os.path.join(REF, path)) as f:
        return textwrap.dedent("".join(f.read().splitlines(keepends=True)[first - 1:last]))


out = {}
# ---- f3 -------------------------------------------------------------------------------------------------------------
g = np.load(os.path.join(HERE, "c1_small.npz"))
segs = [0, 3]
out["interp_segments"] = np.array(segs)
for k, s in enumerate(segs):
    a1, b1, a2, b2 = g["off1"][s], g["off1"][s + 1], g["off2"][s], g["off2"][s + 1]
    x1, x2, T21, T12 = g["xyz1"][a1:b1], g["xyz2"][a2:b2], g["T21r"][a2:b2], g["T12r"][a1:b1]
    for alpha in (0.3, 0.5, 1.0):
        asfloat = alpha == 0.3                                # what np.loadtxt("T21_i.csv") hands the reference
        r = IP.interpolation_segment_level(x1.copy(), x2.copy(), T21.astype(float) if asfloat else T21.copy(),
                                           T12.astype(float) if asfloat else T12.copy(), alpha)
        out[f"interp_{k}_{alpha}"] = np.asarray(r, np.float64)

# ---- f4: labelling --------------------------------------------------------------------------------------------------
rng = np.random.default_rng(404)
n_nodes, n_pts = 180, 6000
tt = np.sort(rng.uniform(0, 1, n_nodes))
nodes = np.stack([40 * tt, 6 * np.sin(7 * tt), 5 * np.cos(5 * tt)], 1) + rng.normal(0, 0.02, (n_nodes, 3))
seg_id = np.floor(tt * 6).astype(float)                       # six segments along the skeleton
xyz_skeleton4 = np.column_stack([nodes, seg_id])
pts = nodes[rng.integers(0, n_nodes, n_pts)] + rng.normal(0, 0.8, (n_pts, 3))
pts[:40] = (nodes[10:50] + nodes[11:51]) / 2                   # exact midpoints: distance ties between neighbouring nodes
ns = dict(np=np, KDTree=KDTree, xyz_skeleton=xyz_skeleton4.copy(), pcd=types.SimpleNamespace(points=pts.copy()))
exec(ref_lines("segment_matching/segment_matching_utils.py", 130, 137), ns)
out.update(label_skeleton=xyz_skeleton4, label_pts=pts, label_nn=np.asarray(ns["nn"], np.int64), label_seg=np.asarray(ns["seg_pcd"]))

# ---- f4: landmark sampling ------------------------------------------------------------------------------------------
code = ref_lines("segment_matching/segment_matching_ml.py", 210, 239)
cases = []
for ci, (n1, n2, vox, method) in enumerate([(60, 75, 0.4, "end"), (33, 20, 0.4, "end"), (90, 64, 0.35, "uniform"),
                                            (12, 12, 1.0, "end"), (5, 40, 0.4, "uniform"), (2, 3, 0.4, "end")]):
    u1, u2 = np.sort(rng.uniform(0, 1, n1)), np.sort(rng.uniform(0, 1, n2))
    mk = lambda u, L: np.stack([L * u, 3 * np.sin(4 * u), 2 * np.cos(3 * u)], 1)
    s1, s2 = mk(u1, 30.0), mk(u2, 33.0) + 0.5
    ns = dict(np=np, options={"pcd_clean_option": {"voxel_size": vox}}, sample_method=method, skel_1=s1.copy(), skel_2=s2.copy(),
              skel_lm1=[], skel_lm2=[])
    exec(code, ns)
    out[f"lm_{ci}_skel1"], out[f"lm_{ci}_skel2"] = s1, s2
    out[f"lm_{ci}_par"] = np.array([vox, 0.0 if method == "end" else 1.0])
    out[f"lm_{ci}_lm1"], out[f"lm_{ci}_lm2"] = np.asarray(ns["skel_lm1"][0]), np.asarray(ns["skel_lm2"][0])
    cases.append(ci)
out["lm_cases"] = np.array(cases)
np.savez_compressed(os.path.join(HERE, "next.npz"), **out)
print("next.npz:", {k: v.shape for k, v in out.items() if k.startswith(("interp_0_0.3", "label_n", "lm_0_lm"))},
      f"{os.path.getsize(os.path.join(HERE, 'next.npz')) / 1e3:.0f} kB")
