"""Writes tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on seeded
synthetic inputs.  Authoring-container only (needs /root/reference); the fixtures it writes are
committed so the GPU box and CI never need the reference tree.

    python tests/golden/make_golden.py

Recorded per case: the inputs, and for both directions (21: mesh2 -> mesh1, 12: mesh1 -> mesh2)
the dense map returned by get_p2p_mapping_from_segments_using_icp, and per landmark partition the
(R, t) my_icp returned, whether it ran, and its number of NN passes.  Also the outputs of the two
functions that follow the hot path in the reference's per-segment body
(double_direction_refinement, evaluate_p2p_matching) for the 'next' rows of SURVEY.md §8f.
"""
import copy
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import reference_loader  # noqa: E402

ref = reference_loader.load()
syn = importlib.import_module("4d_plant_analysis_b200.synthetic")
from sklearn.neighbors import KDTree  # noqa: E402  (same call the reference makes)

_calls = []
_nn_passes = [0]
_orig_icp = ref.RI.my_icp
_orig_nn = ref.RI.get_NN_in_target


def _count_nn(a, b):
    _nn_passes[0] += 1
    return _orig_nn(a, b)


def _rec_icp(a, b):
    _nn_passes[0] = 0
    T, R, t = _orig_icp(a, b)
    _calls.append((R.copy(), np.asarray(t).reshape(3).copy(), _nn_passes[0]))
    return T, R, t


ref.RI.get_NN_in_target = _count_nn      # my_icp resolves it as a module global at call time
ref.SI.my_icp = _rec_icp                 # ditto for the segment function


def run_direction(xyz_t, xyz_s, lm_t, lm_s):
    """mesh1 = target, mesh2 = source, exactly as p2p_matching_seg_icp.py:173."""
    _calls.clear()
    T = ref.SI.get_p2p_mapping_from_segments_using_icp(ref.Mesh(xyz_t.copy()), ref.Mesh(xyz_s.copy()),
                                                       lm_t.copy(), lm_s.copy(), 0.4)
    K = len(lm_t)
    R = np.full((K, 3, 3), np.nan)
    t = np.full((K, 3), np.nan)
    it = np.zeros(K, np.int64)
    valid = np.zeros(K, bool)
    if K > 1:
        n_t = KDTree(lm_t).query(xyz_t, k=2)[1][:, 0]
        n_s = KDTree(lm_s).query(xyz_s, k=2)[1][:, 0]
        valid = np.array([(n_t == i).any() and (n_s == i).any() for i in range(K)])
    else:
        valid[:] = True
    assert valid.sum() == len(_calls), (valid.sum(), len(_calls))
    for i, c in zip(np.nonzero(valid)[0], _calls):
        R[i], t[i], it[i] = c
    return np.asarray(T, np.int64), R, t, it, valid


def run_case(name, batch, post=True):
    out = dict(xyz1=batch.xyz1, xyz2=batch.xyz2, off1=batch.off1, off2=batch.off2,
               lm1=batch.lm1, lm2=batch.lm2, lm_off=batch.lm_off)
    acc = {k: [] for k in ["T21", "T12", "R21", "t21", "it21", "v21", "R12", "t12", "it12", "v12",
                           "T21r", "T12r", "score", "score_raw"]}
    for s in range(batch.n_segments):
        x1, x2, l1, l2 = batch.segment(s)
        T21, R, t, it, v = run_direction(x1, x2, l1, l2)
        acc["T21"].append(T21); acc["R21"].append(R); acc["t21"].append(t); acc["it21"].append(it); acc["v21"].append(v)
        T12, R, t, it, v = run_direction(x2, x1, l2, l1)
        acc["T12"].append(T12); acc["R12"].append(R); acc["t12"].append(t); acc["it12"].append(it); acc["v12"].append(v)
        if post:
            m1, m2 = ref.Mesh(x1.copy()), ref.Mesh(x2.copy())
            acc["score_raw"].append(ref.EV.evaluate_p2p_matching(m1, m2, l1, l2, T21.copy(), T12.copy(),
                                                                 verbose=False, sample_proportion=0.5))
            T21r, T12r = ref.FM.double_direction_refinement(m1, m2, T21.copy(), T12.copy(), None, None)
            acc["T21r"].append(np.asarray(T21r, np.int64)); acc["T12r"].append(np.asarray(T12r, np.int64))
            acc["score"].append(ref.EV.evaluate_p2p_matching(m1, m2, l1, l2, T21r, T12r, verbose=False,
                                                             sample_proportion=0.5))
    for k, v in acc.items():
        if v:
            out[k] = np.concatenate(v) if k not in ("score", "score_raw") else np.asarray(v, np.float64)
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: S={batch.n_segments} N1={batch.off1[-1]} N2={batch.off2[-1]} sumK={batch.lm_off[-1]} "
          f"icp21={int(out['v21'].sum())} iters mean={out['it21'][out['v21']].mean():.2f} "
          f"-> {os.path.getsize(path) / 1e3:.0f} kB")


def tiny_batch():
    """Edge cases: very small partitions, single points, a 2-landmark segment, unequal sizes."""
    rng = np.random.default_rng(77)
    segs = []
    for n1, n2, K in [(5, 7, 2), (40, 3, 2), (12, 12, 3), (1, 9, 1), (30, 31, 1), (64, 90, 4)]:
        L = 3.5 * max(K, 1)
        s1, s2 = np.sort(rng.uniform(0, L, n1)), np.sort(rng.uniform(0, L, n2))
        mk = lambda s: np.stack([s, np.cos(3 * s) + rng.normal(0, .05, len(s)),
                                 np.sin(3 * s) + rng.normal(0, .05, len(s))], 1)
        p1, p2 = mk(s1), mk(s2) + np.array([0.3, -0.2, 0.1])
        lm = np.stack([np.linspace(0, L, K) if K > 1 else np.array([L / 2]), np.zeros(K), np.zeros(K)], 1)
        segs.append((p1 + 20.0, p2 + 20.0, lm + 20.0, lm + 20.0 + np.array([0.3, -0.2, 0.1])))
    return syn.batch_from_segments(segs, name="tiny")


def unit_vectors():
    """Direct calls of the three registration_icp.py functions."""
    rng = np.random.default_rng(5)
    out = {}
    x1, x2, _, _ = syn.make_segment_pair(np.random.default_rng(11), kind="tomato", n1=650, L=40.0)
    _nn_passes[0] = 0
    T, R, t = _orig_icp(x2, x1)
    out.update(icp_src=x2, icp_tgt=x1, icp_T=T, icp_R=R, icp_t=t)
    d, i = _orig_nn(x2, x1)
    out.update(nn_d=d, nn_i=i)
    A = rng.normal(size=(60, 3)) * [5, 2, 0.7]
    Rt = np.linalg.qr(rng.normal(size=(3, 3)))[0]
    Rt *= np.sign(np.linalg.det(Rt))
    B = A @ Rt.T + [1.0, -2.0, 0.5] + rng.normal(size=A.shape) * 0.01
    T, R, t = ref.RI.get_best_fit_transform(A, B)
    out.update(kab_A=A, kab_B=B, kab_T=T, kab_R=R, kab_t=t)
    T, R, t = ref.RI.get_best_fit_transform(A, -B)         # point reflection: det(VU^T) < 0
    out.update(kab3_A=A, kab3_B=-B, kab3_R=R, kab3_t=t)
    ang = np.array([ref.SI.get_triangle_angles(3.0, 2.0, 2.5), ref.SI.get_triangle_angles(1.0, 1.0, 1.0),
                    ref.SI.get_triangle_angles(4.0, 1.5, 3.9)])
    out.update(tri_in=np.array([[3.0, 2.0, 2.5], [1.0, 1.0, 1.0], [4.0, 1.5, 3.9]]), tri_out=ang)
    np.savez_compressed(os.path.join(HERE, "unit.npz"), **out)
    print("unit: ok")


if __name__ == "__main__":
    np.random.seed(0)
    unit_vectors()
    run_case("tiny", tiny_batch(), post=False)
    run_case("c1_small", syn.make_config(1, scale=0.2))
    run_case("c2_small", syn.make_config(2, scale=1 / 12), post=False)
    run_case("c3_small", syn.make_config(3, scale=0.12))
    run_case("c5_small", syn.make_config(5, scale=0.003))
