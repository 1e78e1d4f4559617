"""Pins oracle/local_icp_oracle.py to the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from conftest import golden_segments, load_golden
from oracle import local_icp_oracle as O

CASES = ["tiny", "c1_small", "c2_small", "c3_small", "c5_small"]


def _check_direction(xt, xs, lt, ls, T_ref, R_ref, t_ref, it_ref, v_ref, nn, loop, exact_map):
    det = {}
    T = O.segment_map(xt, xs, lt, ls, nn=nn, loop=loop, details=det)
    assert T.dtype == np.int64 and T.shape == (len(xs),)
    np.testing.assert_array_equal(det["valid"], v_ref)
    np.testing.assert_array_equal(det["iters"][v_ref], it_ref[v_ref])
    np.testing.assert_allclose(det["R"][v_ref], R_ref[v_ref], rtol=0, atol=1e-12)
    np.testing.assert_allclose(det["t"][v_ref], t_ref[v_ref], rtol=0, atol=1e-10)
    if exact_map:
        np.testing.assert_array_equal(T, T_ref)
    else:
        # exact fp64 ties may resolve differently between KD-tree traversal order and lowest index
        bad = np.nonzero(T != T_ref)[0]
        x2t = det["x2t"]
        for i in bad:
            da = np.linalg.norm(x2t[i] - xt[T[i]])
            db = np.linalg.norm(x2t[i] - xt[T_ref[i]])
            assert abs(da - db) <= 1e-6 * max(da, db)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("nn,loop", [("kdtree", True), ("kdtree", False), ("brute", False)])
def test_segment_map_matches_reference(case, nn, loop):
    g = load_golden(case)
    if loop and case in ("c2_small", "c3_small"):
        pytest.skip("per-point loop variant is checked on the smaller cases")
    for s, x1, x2, l1, l2 in golden_segments(g):
        a, b = g["off1"][s], g["off1"][s + 1]
        c, d = g["off2"][s], g["off2"][s + 1]
        e, f = g["lm_off"][s], g["lm_off"][s + 1]
        _check_direction(x1, x2, l1, l2, g["T21"][c:d], g["R21"][e:f], g["t21"][e:f], g["it21"][e:f],
                         g["v21"][e:f], nn, loop, exact_map=(nn == "kdtree"))
        _check_direction(x2, x1, l2, l1, g["T12"][a:b], g["R12"][e:f], g["t12"][e:f], g["it12"][e:f],
                         g["v12"][e:f], nn, loop, exact_map=(nn == "kdtree"))


def test_unit_vectors():
    u = load_golden("unit")
    for nn in ("kdtree", "brute"):
        T, R, t = O.icp(u["icp_src"], u["icp_tgt"], nn=nn)
        np.testing.assert_allclose(R, u["icp_R"], atol=1e-12)
        np.testing.assert_allclose(t, u["icp_t"], atol=1e-10)
        np.testing.assert_allclose(T, u["icp_T"], atol=1e-10)
        assert t.shape == (3, 1)
        d, i = O.nn_in_target(u["icp_src"], u["icp_tgt"], nn=nn)
        assert d.shape == (len(u["icp_src"]), 1) and i.shape == (len(u["icp_src"]),)
        np.testing.assert_array_equal(i, u["nn_i"])
        np.testing.assert_allclose(d, u["nn_d"], rtol=1e-15)
    T, R, t = O.best_fit_transform(u["kab_A"], u["kab_B"])
    np.testing.assert_allclose(R, u["kab_R"], atol=1e-14)
    np.testing.assert_allclose(t, u["kab_t"], atol=1e-13)
    np.testing.assert_allclose(T, u["kab_T"], atol=1e-13)
    assert t.shape == (3,)
    T, R, t = O.best_fit_transform(u["kab3_A"], u["kab3_B"])
    np.testing.assert_allclose(R, u["kab3_R"], atol=1e-14)
    assert np.linalg.det(R) > 0
    for (a, b, c), ang in zip(u["tri_in"], u["tri_out"]):
        np.testing.assert_allclose(O.triangle_angles(a, b, c), ang, rtol=1e-15)


def test_fill_order():
    """p2p_matching_seg_icp.py:422-437: i-1 first (cascading forward), then i+1, then identity/mean shift."""
    I2 = 2 * np.eye(3)
    R = [None, None, I2, None, None]
    t = [None, None, np.ones((3, 1)), None, None]
    x1, x2 = np.ones((4, 3)) * 5, np.ones((3, 3))
    R, t = O.fill_missing(R, t, x1, x2)
    np.testing.assert_array_equal(R[0], np.eye(3))          # 0: no i-1, i+1 missing -> identity
    np.testing.assert_array_equal(t[0].reshape(3), [4, 4, 4])
    np.testing.assert_array_equal(R[1], np.eye(3))          # 1: takes the already filled 0
    np.testing.assert_array_equal(R[3], I2)
    np.testing.assert_array_equal(R[4], I2)


@pytest.mark.parametrize("case", ["c1_small", "c3_small", "c5_small"])
def test_double_direction_refinement_matches_reference(case):
    """the restatement with inverse-index tables reproduces the reference's refined maps"""
    g = load_golden(case)
    for s, x1, x2, l1, l2 in golden_segments(g):
        a, b = g["off1"][s], g["off1"][s + 1]
        c, d = g["off2"][s], g["off2"][s + 1]
        T = O.double_direction_refinement(x1, x2, g["T21"][c:d], g["T12"][a:b], nn="kdtree")
        np.testing.assert_array_equal(T, g["T21r"][c:d])
        np.testing.assert_array_equal(g["T12"][a:b], g["T12r"][a:b])


# ---- rows either side of the hot path, pinned by tests/golden/next.npz -----------------------
def _interp_cases():
    g, nx = load_golden("c1_small"), load_golden("next")
    for k, s in enumerate(nx["interp_segments"]):
        a1, b1, a2, b2 = g["off1"][s], g["off1"][s + 1], g["off2"][s], g["off2"][s + 1]
        for alpha in (0.3, 0.5, 1.0):
            yield g["xyz1"][a1:b1], g["xyz2"][a2:b2], g["T21r"][a2:b2], g["T12r"][a1:b1], alpha, nx[f"interp_{k}_{alpha}"]


def test_interpolation_segment_level_matches_reference():
    n = 0
    for x1, x2, T21, T12, alpha, want in _interp_cases():
        got = O.interpolation_segment_level(x1, x2, T21.astype(float), T12, alpha)      # float maps: what np.loadtxt yields
        assert got.shape == (len(x1) + len(x2), 3)
        np.testing.assert_array_equal(got, want)
        n += 1
    assert n == 6


@pytest.mark.parametrize("nn", ["kdtree", "brute"])
def test_label_points_by_skeleton_matches_reference(nn):
    nx = load_golden("next")
    seg, idx = O.label_points_by_skeleton(nx["label_pts"], nx["label_skeleton"], nn=nn)
    d_got = np.linalg.norm(nx["label_pts"] - nx["label_skeleton"][idx, :3], axis=1)
    d_ref = np.linalg.norm(nx["label_pts"] - nx["label_skeleton"][nx["label_nn"], :3], axis=1)
    diff = idx != nx["label_nn"]
    assert (np.abs(d_got - d_ref)[diff] <= 1e-12 * d_ref[diff]).all()                    # only exact-tie midpoints may differ
    assert diff.sum() <= 40
    np.testing.assert_array_equal(seg[~diff], nx["label_seg"][~diff])


def test_skeleton_landmarks_match_reference():
    nx = load_golden("next")
    for ci in nx["lm_cases"]:
        vox, method = nx[f"lm_{ci}_par"]
        lm1, lm2 = O.skeleton_landmarks(nx[f"lm_{ci}_skel1"], nx[f"lm_{ci}_skel2"], vox, "end" if method == 0 else "uniform")
        np.testing.assert_array_equal(lm1, nx[f"lm_{ci}_lm1"])
        np.testing.assert_array_equal(lm2, nx[f"lm_{ci}_lm2"])
        assert len(lm1) == len(lm2)


def test_interpolation_and_landmark_sampling_properties():
    """size-independent properties of the two restatements: alpha = 0 / 1 reproduce the matched / the original clouds,
    the rows are affine in alpha; the sampled landmarks of a pair have equal counts, start at the first skeleton node and
    are spaced by more than the thinning distance."""
    rng = np.random.default_rng(9)
    for n1, n2 in ((1, 1), (7, 40), (300, 280)):
        x1, x2 = rng.normal(size=(n1, 3)), rng.normal(size=(n2, 3))
        T21, T12 = rng.integers(0, n1, n2), rng.integers(0, n2, n1)
        a0, a1 = O.interpolation_segment_level(x1, x2, T21, T12, 0.0), O.interpolation_segment_level(x1, x2, T21, T12, 1.0)
        np.testing.assert_array_equal(a0, np.vstack([x1[T21], x1]))
        np.testing.assert_array_equal(a1, np.vstack([x2, x2[T12]]))
        np.testing.assert_allclose(O.interpolation_segment_level(x1, x2, T21, T12, 0.25), 0.75 * a0 + 0.25 * a1, rtol=0, atol=1e-15)
    for n1, n2, vox, method in ((50, 70, 0.4, "end"), (9, 200, 0.3, "uniform"), (1, 5, 0.4, "end"), (40, 40, 2.0, "uniform")):
        mk = lambda n: np.cumsum(np.abs(rng.normal(0.5, 0.2, (n, 3))), axis=0)
        s1, s2 = mk(n1), mk(n2)
        lm1, lm2 = O.skeleton_landmarks(s1, s2, vox, method)
        assert len(lm1) == len(lm2) >= 1
        np.testing.assert_array_equal(lm1[0], s1[0]); np.testing.assert_array_equal(lm2[0], s2[0])
        if method == "uniform" and len(lm1) > 1:
            assert (np.linalg.norm(np.diff(lm1, axis=0), axis=1) > 2 * vox).all()
