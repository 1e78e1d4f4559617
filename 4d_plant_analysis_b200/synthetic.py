"""Seeded synthetic plant segments for the five BASELINE.json configs.

Nothing here is on the product path: it only manufactures inputs with the shape of what the
reference's coarse matching hands to the hot path — per matched segment pair two vertex sets
(``mesh1.vertices``, ``mesh2.vertices``) and two equally long landmark polylines (``lm1``, ``lm2``),
see ``p2p_matching_in_organ/p2p_matching_seg_icp.py:152-176`` and the landmark thinning in
``segment_matching/segment_matching_ml.py:210-229`` of the reference.

Everything is generated in float64 with ``np.random.default_rng(1000 * config + frame_index)``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np

VOXEL = {"tomato": 0.8, "maize": 0.8, "arabidopsis": 0.7}


@dataclass
class SegmentPairBatch:
    """CSR batch of matched segment pairs (one or many frame pairs flattened).

    xyz1/xyz2: (sum N1, 3) / (sum N2, 3) float64 vertices of frame-1 / frame-2 segments
    off1/off2: (S+1,) int64 CSR offsets; lm1/lm2: (sum K, 3) float64; lm_off: (S+1,) int64
    frame_pair: (S,) int32 frame-pair id of each segment pair (for frame-pair/s accounting)
    """

    xyz1: np.ndarray
    xyz2: np.ndarray
    off1: np.ndarray
    off2: np.ndarray
    lm1: np.ndarray
    lm2: np.ndarray
    lm_off: np.ndarray
    frame_pair: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    name: str = ""

    @property
    def n_segments(self) -> int:
        return len(self.off1) - 1

    @property
    def n_frame_pairs(self) -> int:
        return int(self.frame_pair.max()) + 1 if len(self.frame_pair) else 0

    @property
    def n_correspondences(self) -> int:
        """|T21| + |T12| summed over segment pairs (metric M1 numerator)."""
        return int(self.off1[-1] + self.off2[-1])

    def segment(self, s: int):
        a, b = self.off1[s], self.off1[s + 1]
        c, d = self.off2[s], self.off2[s + 1]
        e, f = self.lm_off[s], self.lm_off[s + 1]
        return self.xyz1[a:b], self.xyz2[c:d], self.lm1[e:f], self.lm2[e:f]

    def segments(self):
        return [self.segment(s) for s in range(self.n_segments)]

    def take(self, idx: Sequence[int]) -> "SegmentPairBatch":
        return batch_from_segments([self.segment(int(s)) for s in idx],
                                   frame_pair=self.frame_pair[np.asarray(idx, dtype=np.int64)],
                                   name=self.name)


def batch_from_segments(segs, frame_pair=None, name="") -> SegmentPairBatch:
    S = len(segs)
    n1 = np.array([len(s[0]) for s in segs], dtype=np.int64)
    n2 = np.array([len(s[1]) for s in segs], dtype=np.int64)
    k = np.array([len(s[2]) for s in segs], dtype=np.int64)
    for s in segs:
        if len(s[2]) != len(s[3]):
            raise ValueError("lm1 and lm2 must have the same number of landmarks")
    cat = lambda i: (np.ascontiguousarray(np.concatenate([np.asarray(s[i], np.float64).reshape(-1, 3)
                                                          for s in segs], axis=0))
                     if S else np.zeros((0, 3)))
    z = lambda n: np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    fp = np.zeros(S, np.int32) if frame_pair is None else np.asarray(frame_pair, np.int32)
    return SegmentPairBatch(cat(0), cat(1), z(n1), z(n2), cat(2), cat(3), z(k), fp, name)


def concat_batches(batches: List[SegmentPairBatch], name="") -> SegmentPairBatch:
    segs, fp, base = [], [], 0
    for b in batches:
        segs += b.segments()
        fp.append(b.frame_pair + base)
        base += max(b.n_frame_pairs, 1)
    return batch_from_segments(segs, np.concatenate(fp) if fp else None, name)


# ------------------------------------------------------------------------------------------------
# geometry of one segment
# ------------------------------------------------------------------------------------------------
def _rot(axis, ang):
    axis = axis / np.linalg.norm(axis)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def _arc(s, kappa):
    """Planar arc of curvature kappa in the local (x, z) plane, unit speed; returns centre line,
    normal and binormal frames for arc-length parameters s."""
    s = np.asarray(s, np.float64)
    if abs(kappa) < 1e-9:
        c = np.stack([s, np.zeros_like(s), np.zeros_like(s)], 1)
        n = np.tile([0.0, 0.0, 1.0], (len(s), 1))
    else:
        c = np.stack([np.sin(kappa * s) / kappa, np.zeros_like(s), (1 - np.cos(kappa * s)) / kappa], 1)
        n = np.stack([-np.sin(kappa * s), np.zeros_like(s), np.cos(kappa * s)], 1)
    b = np.tile([0.0, 1.0, 0.0], (len(s), 1))
    return c, n, b


def _surface(rng, n, L, kappa, shape, size, jitter=0.05):
    s = rng.uniform(0.0, L, n)
    c, nn, bb = _arc(s, kappa)
    if shape == "tube":
        th = rng.uniform(0.0, 2 * np.pi, n)
        rad = np.maximum(size + rng.normal(0.0, jitter, n), 0.3)
        p = c + rad[:, None] * (np.cos(th)[:, None] * nn + np.sin(th)[:, None] * bb)
    else:  # ribbon of width `size`: offset along the binormal, thin along the normal
        u = rng.uniform(0.3, size / 2, n) * rng.choice([-1.0, 1.0], n)
        p = c + u[:, None] * bb + rng.normal(0.0, jitter, n)[:, None] * nn
    return p


def _thin(nodes, gap):
    keep = [nodes[0]]
    for x in nodes:
        if np.linalg.norm(keep[-1] - x) > gap:
            keep.append(x)
    return np.vstack(keep)


def _landmarks(rng, L1, k1, L2, k2, voxel, mirrored):
    """Mirrors segment_matching_ml.py:210-229: skeleton nodes thinned to > 2*voxel (3*voxel in the
    mirrored branch), both truncated to the same length, every second node kept plus the last."""
    gap = (3 if mirrored else 2) * voxel
    out = []
    for L, kap in ((L1, k1), (L2, k2)):
        sn = np.arange(0.0, L + 1e-9, 0.5 * voxel)
        c, _, _ = _arc(sn, kap)
        c = c + rng.normal(0.0, 0.03, c.shape)
        out.append(_thin(c, gap))
    keep = min(len(out[0]), len(out[1]))
    res = []
    for sk in out:
        sk = sk[:keep]
        res.append(np.vstack([sk[0:len(sk) - 1:2], sk[-1]]))
    return res[0], res[1]


def make_segment_pair(rng, *, kind="tomato", n1=None, L=None, force_k1=False, empty_partition=False):
    voxel = VOXEL[kind]
    shape = "ribbon" if kind == "maize" else "tube"
    if kind == "maize":
        L = rng.uniform(60, 150) if L is None else L
        size = rng.uniform(4, 8)
        n1 = int(rng.uniform(3200, 4800)) if n1 is None else n1
    else:
        if L is None:
            L = rng.uniform(15, 60) if kind == "tomato" else rng.uniform(5, 30)
        size = rng.uniform(0.8, 2.0)
        if n1 is None:
            n1 = int(rng.uniform(300, 1100)) if kind == "tomato" else int(rng.uniform(150, 1000))
    kap1 = rng.uniform(-0.02, 0.02)
    grow = rng.uniform(1.02, 1.12)
    L2, kap2 = L * grow, kap1 + rng.uniform(-0.005, 0.005)
    n2 = int(round(n1 * grow))
    p1 = _surface(rng, n1, L, kap1, shape, size)
    p2 = _surface(rng, n2, L2, kap2, shape, size * rng.uniform(1.0, 1.05))
    if force_k1:
        lm1 = _arc([L / 2], kap1)[0]
        lm2 = _arc([L2 / 2], kap2)[0]
    else:
        lm1, lm2 = _landmarks(rng, L, kap1, L2, kap2, voxel, mirrored=bool(rng.random() < 0.3))
        if empty_partition and len(lm1) > 2:
            # a landmark pushed off the surface so that no vertex of one frame picks it first
            j = int(rng.integers(1, len(lm1) - 1))
            (lm1 if rng.random() < 0.5 else lm2)[j] += np.array([0.0, 0.0, 40.0])
    # frame-2 motion about the base: rotation <= 8 deg, translation <= 1.5
    Rm = _rot(rng.normal(size=3), np.deg2rad(rng.uniform(0, 8)))
    tm = rng.uniform(-1.5, 1.5, 3) * (1 / np.sqrt(3))
    p2, lm2 = p2 @ Rm.T + tm, lm2 @ Rm.T + tm
    # place in the plant frame (coordinates up to +-100)
    Rp = _rot(rng.normal(size=3), rng.uniform(0, np.pi))
    tp = rng.uniform(-60, 60, 3)
    f = lambda a: np.ascontiguousarray(a @ Rp.T + tp)
    return f(p1), f(p2), f(lm1), f(lm2)


# ------------------------------------------------------------------------------------------------
# the five configs
# ------------------------------------------------------------------------------------------------
def _frame_pair(config, frame_index, S, kind, **kw):
    rng = np.random.default_rng(1000 * config + frame_index)
    segs = []
    for _ in range(S):
        extra = {}
        if config == 3:
            u = rng.random()
            extra = dict(force_k1=u < 0.05, empty_partition=0.05 <= u < 0.10)
        segs.append(make_segment_pair(rng, kind=kind, **extra, **kw))
    return segs


def make_config(config: int, *, scale: float = 1.0, frame_pairs: int | None = None,
                frame_offset: int = 0) -> SegmentPairBatch:
    """config 1..5 = BASELINE.json configs[0..4]. ``scale`` < 1 shrinks the number of segment pairs
    (parity-sized cases); ``frame_pairs`` overrides the number of frame pairs (C4: 15);
    ``frame_offset`` shifts the frame index that seeds each frame pair (distinct data per rank)."""
    if config == 1:      # tomato: ~20k pts/frame, ~30 segments
        fps, S, kind, kw = frame_pairs or 1, 30, "tomato", {}
    elif config == 2:    # maize: ~50k pts/frame, 12 long ribbons
        fps, S, kind, kw = frame_pairs or 1, 12, "maize", {}
    elif config == 3:    # arabidopsis: ~100k pts/frame, ~200 small segments, K==1 and empty partitions
        fps, S, kind, kw = frame_pairs or 1, 200, "arabidopsis", {}
    elif config == 4:    # 16-frame tomato series = 15 consecutive pairs
        fps, S, kind, kw = frame_pairs or 15, 30, "tomato", {}
    elif config == 5:    # 1M points, 2000 segment pairs
        fps, S, kind, kw = frame_pairs or 1, 2000, "tomato", dict(L=None)
    else:
        raise ValueError(f"unknown config {config}")
    S = max(1, int(round(S * scale)))
    segs, fp = [], []
    for f0 in range(fps):
        f = f0 + frame_offset
        if config == 5:
            rng = np.random.default_rng(1000 * config + f)
            these = [make_segment_pair(rng, kind="tomato", n1=int(rng.uniform(400, 600)),
                                       L=rng.uniform(15, 25)) for _ in range(S)]
        else:
            these = _frame_pair(config, f, S, kind, **kw)
        segs += these
        fp += [f0] * len(these)
    return batch_from_segments(segs, np.asarray(fp, np.int32), name=f"C{config}")
