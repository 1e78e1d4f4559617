"""CPU oracle for the local-ICP point-to-point matching hot path.   *** TEST INFRASTRUCTURE ***

This is a float64 numpy restatement of the reference algorithm, used ONLY as the checker in
``tests/``, ``__graft_entry__.smoke()`` and as the timed CPU arm of ``bench.py``
(``cpu_baseline`` / ``--impl reference``).  Nothing in the product package imports it.

Parity pin: the reference ships no tests or golden vectors for this path, so the
pin is ``tests/golden/*.npz`` — outputs of the UNMODIFIED reference functions run in the authoring
container by ``tests/golden/make_golden.py`` — which ``tests/test_oracle.py`` replays against this
file (maps identical, R/t to 1e-12, iteration counts identical).

Reference functions restated (paths relative to the reference root):
  nn_in_target         <- registration/registration_icp.py:49-57   (get_NN_in_target)
  best_fit_transform   <- registration/registration_icp.py:8-46    (get_best_fit_transform)
  icp                  <- registration/registration_icp.py:60-98   (my_icp)
  triangle_angles      <- p2p_matching_in_organ/p2p_matching_seg_icp.py:27-34
  segment_map          <- p2p_matching_in_organ/p2p_matching_seg_icp.py:383-489
                          (get_p2p_mapping_from_segments_using_icp)
  double_direction_refinement <- p2p_matching_in_organ/p2p_matching_fm.py:75-132
  interpolation_segment_level <- interpolation/interpolation_pcd.py:122-134
  label_points_by_skeleton    <- segment_matching/segment_matching_utils.py:130-137
  skeleton_landmarks          <- segment_matching/segment_matching_ml.py:210-239
                          (the last three pinned by tests/golden/next.npz, written by make_golden_next.py)

Third-party arithmetic on the path that is not in the reference tree:
  * scikit-learn ``sklearn.neighbors.KDTree`` (reference pins scikit-learn~=0.24.2, requirements.txt:11;
    this image has 1.9.0).  Exact Euclidean k-NN; reduced distance sum((x-y)**2) accumulated
    dimension by dimension in float64, sqrt at the end; leaf_size 40 => a target of <= 80 points is a
    single leaf scanned in index order with a strict '<', i.e. lowest index wins ties.  Larger
    targets: first point met in depth-first, closer-child-first order wins an exact tie.
    ``nn="kdtree"`` calls the same library the reference calls; ``nn="brute"`` restates the
    published semantics (exact NN, lowest index on ties) without it.
  * numpy ``np.linalg.svd`` (LAPACK gesdd) and ``np.linalg.det`` for the 3x3 Kabsch solve.
"""
from __future__ import annotations

import numpy as np

MAX_ITER = 500           # registration_icp.py:79
COST_RISE_TOL = 1e-4     # registration_icp.py:87
COST_FLAT_TOL = 0.01     # registration_icp.py:87
OBTUSE_EPS = 1e-5        # p2p_matching_seg_icp.py:471
COS_DEN_EPS = 1e-5       # p2p_matching_seg_icp.py:30-32


# ------------------------------------------------------------------------------------------------
# nearest neighbours
# ------------------------------------------------------------------------------------------------
def _knn_brute(query, target, k):
    """Exact k-NN, float64, difference form accumulated x then y then z (sklearn
    _dist_metrics rdist), stable ordering so the lowest index wins ties."""
    query = np.asarray(query, np.float64).reshape(-1, 3)
    target = np.asarray(target, np.float64).reshape(-1, 3)
    n, m = len(query), len(target)
    dist = np.empty((n, k))
    ind = np.empty((n, k), np.int64)
    step = max(1, (1 << 22) // max(m, 1))
    for a in range(0, n, step):
        q = query[a:a + step]
        d = (q[:, None, 0] - target[None, :, 0]) ** 2
        d += (q[:, None, 1] - target[None, :, 1]) ** 2
        d += (q[:, None, 2] - target[None, :, 2]) ** 2
        if k == 1:
            j = np.argmin(d, axis=1)[:, None]          # first minimum = lowest index
        else:
            j = np.argsort(d, axis=1, kind="stable")[:, :k]
        ind[a:a + step] = j
        dist[a:a + step] = np.sqrt(np.take_along_axis(d, j, axis=1))
    return dist, ind


def knn(query, target, k=1, nn="kdtree"):
    """(dist (n,k) float64, ind (n,k) int64) like KDTree(target).query(query, k)."""
    if nn == "kdtree":
        from sklearn.neighbors import KDTree
        return KDTree(target).query(query, k=k, return_distance=True)
    if nn == "brute":
        return _knn_brute(query, target, k)
    raise ValueError(nn)


def nn_in_target(xyz1, xyz2, nn="kdtree"):
    """registration_icp.py:49-57 — distance is (n,1), indices (n,)."""
    d, i = knn(xyz1, xyz2, 1, nn)
    return d, i.reshape(-1)


# ------------------------------------------------------------------------------------------------
# Kabsch and ICP
# ------------------------------------------------------------------------------------------------
def best_fit_transform(A, B):
    """registration_icp.py:8-46.  Returns (T 4x4 [R^T top-left, t in the last ROW], R, t(3,))."""
    A = np.asarray(A, np.float64)
    B = np.asarray(B, np.float64)
    assert len(A) == len(B)
    cA, cB = A.mean(axis=0), B.mean(axis=0)
    H = (A - cA).T @ (B - cB)
    U, _, Vt = np.linalg.svd(H)
    R = Vt.T @ U.T
    if np.linalg.det(R) < 0:
        Vt[2, :] *= -1
        R = Vt.T @ U.T
    t = cB - R @ cA
    T = np.identity(4)
    T[0:3, 0:3] = R.T
    T[3, 0:3] = t
    return T, R, t


def icp(src, tgt, nn="kdtree", info=None, reject_dist=np.inf):
    """registration_icp.py:60-98 (my_icp).  Returns (T, R (3,3), t (3,1)); x -> x R^T + t^T maps src on tgt.
    ``info`` (dict), if given, receives 'iters' (number of NN passes) and 'cost' (their mean distances).
    ``reject_dist`` is NOT reference behaviour (the reference rejects nothing, registration_icp.py:85-90): it restates the
    correspondence-reject hook of the CUDA path (lip_params.reject_dist) — pairs farther apart than this are left out of the
    Kabsch fit, the cost stays the mean over ALL pairs, and a pass that keeps no pair applies the identity increment."""
    X = np.array(src, np.float64).reshape(-1, 3).T.copy()
    Y = np.array(tgt, np.float64).reshape(-1, 3).T.copy()
    R_inc = np.identity(3)
    t_inc = (Y.mean(axis=1) - X.mean(axis=1)).reshape(3, 1)
    R = np.identity(3)
    t = np.zeros((3, 1))
    cost = []
    for i in range(MAX_ITER):
        X = R_inc @ X + t_inc
        R = R_inc @ R
        t = R_inc @ t + t_inc
        d, idx = nn_in_target(X.T, Y.T, nn)
        cost.append(np.mean(d))
        if i > 1 and (cost[-1] > cost[-2] + COST_RISE_TOL or abs(cost[-1] - cost[-2]) < COST_FLAT_TOL):
            break
        keep = np.asarray(d).reshape(-1) <= reject_dist
        if keep.all():
            _, R_inc, t_inc = best_fit_transform(X.T, Y.T[idx])
        elif keep.any():
            _, R_inc, t_inc = best_fit_transform(X.T[keep], Y.T[idx][keep])
        else:
            R_inc, t_inc = np.identity(3), np.zeros(3)
        t_inc = t_inc.reshape(3, 1)
    if info is not None:
        info["iters"] = len(cost)
        info["cost"] = np.asarray(cost)
    T = np.identity(4)
    T[0:3, 0:3] = R
    T[3, 0:3] = t[:, 0]
    return T, R, t


# ------------------------------------------------------------------------------------------------
# segment-level map
# ------------------------------------------------------------------------------------------------
def triangle_angles(a, b, c):
    """p2p_matching_seg_icp.py:27-34: angles opposite a, b, c (law of cosines, +1e-5 denominators)."""
    assert a + b >= c and a + c >= b and b + c >= a
    return np.arccos([(b ** 2 + c ** 2 - a ** 2) / (2 * c * b + COS_DEN_EPS),
                      (c ** 2 + a ** 2 - b ** 2) / (2 * a * c + COS_DEN_EPS),
                      (a ** 2 + b ** 2 - c ** 2) / (2 * a * b + COS_DEN_EPS)])


def fill_missing(R_list, t_list, xyz1, xyz2):
    """p2p_matching_seg_icp.py:422-437: ascending pass; missing entry i takes i-1 (already filled),
    else i+1, else (I, mean(xyz1) - mean(xyz2))."""
    K = len(R_list)
    for i in range(K):
        if R_list[i] is None:
            if i > 0 and R_list[i - 1] is not None:
                R_list[i] = R_list[i - 1].copy()
            elif i < K - 1 and R_list[i + 1] is not None:
                R_list[i] = R_list[i + 1].copy()
            else:
                R_list[i] = np.eye(3)
        if t_list[i] is None:
            if i > 0 and t_list[i - 1] is not None:
                t_list[i] = t_list[i - 1].copy()
            elif i < K - 1 and t_list[i + 1] is not None:
                t_list[i] = t_list[i + 1].copy()
            else:
                t_list[i] = (np.mean(xyz1, axis=0) - np.mean(xyz2, axis=0)).reshape(3, 1)
    return R_list, t_list


def blend_weights(lm2, nn2, d2, loop=False):
    """p2p_matching_seg_icp.py:454-479: per-point weights (f0, f1) of the transforms of the two
    nearest landmarks.  loop=True walks the points one by one like the reference (and asserts the
    triangle inequality like it); loop=False is the same arithmetic vectorised (cosines clamped)."""
    n = len(nn2)
    if nn2.shape[1] == 1:
        return np.full(n, 0.5), np.full(n, 0.5)
    f0 = np.empty(n)
    f1 = np.empty(n)
    if loop:
        for i in range(n):
            g = np.linalg.norm(lm2[nn2[i, 1]] - lm2[nn2[i, 0]])
            _, B, C = triangle_angles(g, d2[i, 0], d2[i, 1])
            if C > np.pi / 2 - OBTUSE_EPS:
                f0[i], f1[i] = 1.0, 0.0
            else:
                cb, cc = np.abs(np.cos(B)), np.abs(np.cos(C))
                f0[i], f1[i] = cb / (cb + cc), cc / (cb + cc)
        return f0, f1
    a = np.linalg.norm(lm2[nn2[:, 1]] - lm2[nn2[:, 0]], axis=1)
    b, c = d2[:, 0], d2[:, 1]
    cosB = np.clip((c ** 2 + a ** 2 - b ** 2) / (2 * a * c + COS_DEN_EPS), -1.0, 1.0)
    cosC = np.clip((a ** 2 + b ** 2 - c ** 2) / (2 * a * b + COS_DEN_EPS), -1.0, 1.0)
    obtuse = np.arccos(cosC) > np.pi / 2 - OBTUSE_EPS
    cb, cc = np.abs(np.cos(np.arccos(cosB))), np.abs(np.cos(np.arccos(cosC)))
    with np.errstate(invalid="ignore", divide="ignore"):
        f0 = np.where(obtuse, 1.0, cb / (cb + cc))
        f1 = np.where(obtuse, 0.0, cc / (cb + cc))
    return f0, f1


def segment_map(xyz1, xyz2, lm1, lm2, nn="kdtree", loop=False, details=None, reject_dist=np.inf):
    """p2p_matching_seg_icp.py:383-489.  T21 (N2,) int64: for each vertex of mesh2 the index of its
    match in mesh1.  ``details`` (dict) receives the per-landmark transforms after filling ('R' (K,3,3),
    't' (K,3)), 'valid' (K,) whether an ICP ran, 'iters' (K,) NN passes (0 where none), the landmark
    assignment 'nn1'/'nn2'/'d2' and the blended points 'x2t'."""
    xyz1 = np.array(xyz1, np.float64).reshape(-1, 3)
    xyz2 = np.array(xyz2, np.float64).reshape(-1, 3)
    lm1 = np.asarray(lm1, np.float64).reshape(-1, 3)
    lm2 = np.asarray(lm2, np.float64).reshape(-1, 3)
    K = lm1.shape[0]
    kk = 2 if K > 1 else 1
    _, nn1 = knn(xyz1, lm1, kk, nn)
    d2, nn2 = knn(xyz2, lm2, kk, nn)
    R_list, t_list, iters = [], [], np.zeros(K, np.int64)
    if K > 1:
        for i in range(K):
            tgt = xyz1[nn1[:, 0] == i]
            src = xyz2[nn2[:, 0] == i]
            if len(src) > 0 and len(tgt) > 0:
                inf = {}
                _, R, t = icp(src, tgt, nn, inf, reject_dist)
                iters[i] = inf["iters"]
            else:
                R = t = None
            R_list.append(R)
            t_list.append(t)
        valid = np.array([r is not None for r in R_list])
        R_list, t_list = fill_missing(R_list, t_list, xyz1, xyz2)
    else:
        inf = {}
        _, R, t = icp(xyz2, xyz1, nn, inf, reject_dist)
        iters[0] = inf["iters"]
        R_list, t_list, valid = [R], [t], np.array([True])
    Rs = np.stack(R_list)
    ts = np.stack([np.asarray(t).reshape(3) for t in t_list])
    f0, f1 = blend_weights(lm2, nn2, d2, loop=loop)
    ia, ib = nn2[:, 0], nn2[:, -1]
    xa = np.einsum("nj,nij->ni", xyz2, Rs[ia]) + ts[ia]
    xb = np.einsum("nj,nij->ni", xyz2, Rs[ib]) + ts[ib]
    x2t = f0[:, None] * xa + f1[:, None] * xb
    _, T21 = knn(x2t, xyz1, 1, nn)
    if details is not None:
        details.update(R=Rs, t=ts, valid=valid, iters=iters, nn1=nn1, nn2=nn2, d2=d2, x2t=x2t, f0=f0, f1=f1)
    return T21.reshape(-1).astype(np.int64)


def segment_pair_maps(xyz1, xyz2, lm1, lm2, nn="kdtree", loop=False):
    """The pair of calls the reference makes per matched segment pair
    (p2p_matching_seg_icp.py:173-176): T21 maps mesh2 -> mesh1, T12 maps mesh1 -> mesh2."""
    T21 = segment_map(xyz1, xyz2, lm1, lm2, nn, loop)
    T12 = segment_map(xyz2, xyz1, lm2, lm1, nn, loop)
    return T21, T12


# ------------------------------------------------------------------------------------------------
# post-processing of the dense maps: p2p_matching_fm.py:75-132
# ------------------------------------------------------------------------------------------------
def double_direction_refinement(xyz1, xyz2, T21, T12, nn="kdtree"):
    """p2p_matching_in_organ/p2p_matching_fm.py:75-132, restated with inverse-index tables instead of the
    reference's np.unique / np.where inside the loops (same visiting order, same in-place semantics).
    Returns the refined copy of T21 (T12 is not modified by the reference either)."""
    v2 = np.asarray(xyz2, np.float64)
    T21 = np.array(T21, np.int64)
    T12 = np.asarray(T12, np.int64)
    n1, n2 = len(xyz1), len(T21)
    _, nn5 = knn(xyz1, xyz1, min(5, n1), nn)                      # :78-79 (row = [self, 4 nearest])
    # sources of every target, ascending (what np.where(T21 == t)[0] returns), as sorted python lists
    src_of = [[] for _ in range(n1)]
    for i in range(n2):
        src_of[T21[i]].append(i)
    tgt_of = [[] for _ in range(n2)]                              # np.where(T12 == ind)[0]
    for t in range(n1):
        tgt_of[T12[t]].append(t)

    def keep_of(t):                                               # :88-96 / :113-121: smallest cycle error, first on ties
        best, keep = 10 ** 4, src_of[t][0]
        ref = v2[T12[t]]
        for ind in src_of[t]:
            d = np.linalg.norm(v2[ind] - ref)
            if d < best:
                best, keep = d, ind
        return keep

    def move(ind, t_from, t_to):
        src_of[t_from].remove(ind)
        lst = src_of[t_to]
        k = 0
        while k < len(lst) and lst[k] < ind:
            k += 1
        lst.insert(k, ind)
        T21[ind] = t_to

    for t in [t for t in range(n1) if src_of[t]]:                 # np.unique(T21) taken once, :85
        if len(src_of[t]) == 1:
            continue
        keep = keep_of(t)
        for ind in list(src_of[t]):
            if ind == keep:
                continue
            cand = tgt_of[ind] if tgt_of[ind] else list(nn5[t, 1:])   # :100-103
            t_keep = cand[0]
            for c in cand:                                        # the LAST candidate no source maps to, :105-107
                if not src_of[c]:
                    t_keep = c
            move(ind, t, t_keep)
    for t in range(n1):                                           # :111-130
        if src_of[t]:
            continue
        for nb in nn5[t, 1:]:
            if not src_of[nb]:
                continue
            keep = keep_of(nb)
            for ind in list(src_of[nb]):
                if ind != keep:
                    move(ind, nb, t)
    return T21


def interpolation_segment_level(xyz1, xyz2, T21, T12, alpha):
    """interpolation/interpolation_pcd.py:122-134: mesh2 pulled back onto its matches in mesh1, then mesh1 pushed towards
    its matches in mesh2; rows stacked in that order."""
    xyz1, xyz2 = np.asarray(xyz1, np.float64), np.asarray(xyz2, np.float64)
    inter_0 = (1 - alpha) * xyz1[np.asarray(T21).astype(int)] + alpha * xyz2        # :128-129
    inter_1 = (1 - alpha) * xyz1 + alpha * xyz2[np.asarray(T12).astype(int)]        # :131-132
    return np.vstack([inter_0, inter_1])


def label_points_by_skeleton(xyz_pcd, xyz_skeleton, nn="kdtree"):
    """segment_matching/segment_matching_utils.py:130-137: segment id (4th column) of the nearest skeleton node."""
    xyz_skeleton = np.asarray(xyz_skeleton, np.float64)
    seg_skeleton, nodes = xyz_skeleton[:, 3], xyz_skeleton[:, :3]
    _, idx = knn(np.asarray(xyz_pcd, np.float64), nodes, 1, nn)
    idx = idx.reshape(-1)
    return seg_skeleton[idx], idx


def skeleton_landmarks(skel_1, skel_2, voxel_size, sample_method="end"):
    """segment_matching/segment_matching_ml.py:210-239: thin both polylines to a spacing of 2 * voxel_size, then pair
    landmarks either by truncating to the shorter and keeping every second node plus the last ("end"), or by striding."""
    def thin(sk):
        kept = [sk[0]]
        for x in sk:                                          # :212-215 / :217-220
            if np.linalg.norm(kept[-1] - x) > 2 * voxel_size:
                kept.append(x)
        return np.vstack(kept)
    a, b = thin(np.asarray(skel_1, np.float64)), thin(np.asarray(skel_2, np.float64))
    if sample_method == "end":                                # :225-230
        n = min(len(a), len(b))
        a, b = a[:n], b[:n]
        return np.vstack([a[0:len(a) - 1:2], a[-1]]), np.vstack([b[0:len(b) - 1:2], b[-1]])
    half = min(len(a), len(b)) / 2                            # :232-239
    a, b = a[0:len(a):int(len(a) / half) + 1], b[0:len(b):int(len(b) / half) + 1]
    n = min(len(a), len(b))
    return a[:n], b[:n]
