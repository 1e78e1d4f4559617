"""Round-2 golden rows from the UNMODIFIED reference (authoring container only, like make_golden.py):

    python tests/golden/make_golden_r2.py        ->  tests/golden/edge_r2.npz

  seg 0  K == 1, 2,500 x 2,600 vertices: the whole-segment ICP branch (p2p_matching_seg_icp.py:440-449) on a target far larger
         than one search tile — pins the chunked path of icp_big_kernel to reference rows.
  seg 1  K == 2, ~330 vertices per partition on a 0.5 lattice with duplicated points: targets of more than 80 points (sklearn
         builds a real tree there) holding exact float64 distance ties.
  seg 2  K == 4 with thin ends: 2-point and collinear partitions (rank-1 cross-covariance: the reference's rotation is LAPACK's
         arbitrary completion) feeding the blend of their neighbours.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (imports the reference under the stub recipe and installs the recorders)


def batch():
    rng = np.random.default_rng(2026)
    segs = []
    x1, x2, _, _ = G.syn.make_segment_pair(np.random.default_rng(31), kind="maize", n1=2500, L=70.0, force_k1=True)
    c = x1.mean(axis=0)
    segs.append((x1, x2[:2600], c[None, :], (x2.mean(axis=0))[None, :]))
    # lattice tube with duplicates
    L = 8.0
    def lattice(n):
        s = rng.uniform(0, L, n); th = rng.uniform(0, 2 * np.pi, n)
        p = np.stack([s, 1.5 * np.cos(th), 1.5 * np.sin(th)], 1)
        p = np.round(p * 2) / 2
        p[5] = p[4]; p[17] = p[4]; p[100] = p[99]
        return p + 30.0
    lm = np.array([[2.0, 0, 0], [6.0, 0, 0]]) + 30.0
    segs.append((lattice(660), lattice(700) + [0.5, 0.0, 0.0], lm, lm + [0.5, 0.0, 0.0]))
    # thin ends: 2-point and collinear partitions
    K = 4
    lm = np.stack([np.array([1.0, 4.0, 7.0, 10.0]), np.zeros(K), np.zeros(K)], 1) - 15.0
    def thin(shift):
        mid = np.stack([rng.uniform(2.6, 8.4, 160), rng.normal(0, 0.7, 160), rng.normal(0, 0.7, 160)], 1)
        end0 = np.array([[0.4, 0.2, 0.1], [1.3, -0.3, 0.2]])                       # two points
        end1 = np.stack([np.linspace(9.0, 11.0, 6), np.zeros(6), np.zeros(6)], 1)  # collinear, on the landmark line
        return np.vstack([end0, mid, end1]) - 15.0 + shift
    segs.append((thin(0.0), thin(np.array([0.2, 0.1, -0.1])), lm, lm + [0.2, 0.1, -0.1]))
    return G.syn.batch_from_segments(segs, name="edge_r2")


if __name__ == "__main__":
    np.random.seed(0)
    G.run_case("edge_r2", batch(), post=False)
