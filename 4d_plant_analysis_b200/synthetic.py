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


# ------------------------------------------------------------------------------------------------
# time series (BASELINE configs[3]: a growth series, all consecutive frame pairs batched)
# ------------------------------------------------------------------------------------------------
@dataclass
class SeriesBatch:
    """A growth series of one or several plants: every frame's segments are stored ONCE in ``frames``; segment pair s of
    the batch (frame k -> frame k + 1 of one plant) reads mesh1 at ``row1[s]`` and mesh2 at ``row2[s]`` — frame k is day2 of
    frame pair k - 1 and day1 of frame pair k (reference: run_registration_pipeline.py:64,100-104).  off1 / off2 / lm_off /
    lm1 / lm2 / frame_pair describe the batch exactly as in SegmentPairBatch (sizes and the layout of the outputs)."""

    frames: np.ndarray        # (n_rows, 3) float64
    row1: np.ndarray          # (S,) int64
    row2: np.ndarray          # (S,) int64
    off1: np.ndarray
    off2: np.ndarray
    lm1: np.ndarray
    lm2: np.ndarray
    lm_off: np.ndarray
    frame_pair: np.ndarray
    name: str = ""

    @property
    def n_segments(self) -> int:
        return len(self.off1) - 1

    @property
    def n_frame_pairs(self) -> int:
        return int(self.frame_pair.max()) + 1 if len(self.frame_pair) else 0

    @property
    def n_correspondences(self) -> int:
        return int(self.off1[-1] + self.off2[-1])

    def as_batch(self) -> SegmentPairBatch:
        """the same segment pairs with every mesh gathered into xyz1 / xyz2 (what a caller without the series entry point
        would upload: every shared frame twice)"""
        n1, n2 = np.diff(self.off1), np.diff(self.off2)
        x1 = np.concatenate([self.frames[r:r + n] for r, n in zip(self.row1, n1)], axis=0)
        x2 = np.concatenate([self.frames[r:r + n] for r, n in zip(self.row2, n2)], axis=0)
        return SegmentPairBatch(np.ascontiguousarray(x1), np.ascontiguousarray(x2), self.off1, self.off2, self.lm1, self.lm2,
                                self.lm_off, self.frame_pair, self.name)

    def take_frame_pairs(self, ids: Sequence[int]) -> "SeriesBatch":
        """the sub-series holding the given frame pairs (ascending ids): only the meshes they use are kept, each once"""
        ids = np.asarray(sorted(int(i) for i in ids))
        keep = np.nonzero(np.isin(self.frame_pair, ids))[0]
        n1, n2 = np.diff(self.off1)[keep], np.diff(self.off2)[keep]
        # distinct (row, count) meshes in row order
        rows = sorted({(int(r), int(n)) for r, n in zip(self.row1[keep], n1)} | {(int(r), int(n)) for r, n in zip(self.row2[keep], n2)})
        new_row, parts, at = {}, [], 0
        for r, n in rows:
            new_row[r] = at
            parts.append(self.frames[r:r + n])
            at += n
        z = lambda n: np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
        lm_keep = np.concatenate([np.arange(self.lm_off[s], self.lm_off[s + 1]) for s in keep]) if len(keep) else np.zeros(0, np.int64)
        remap = {int(f): i for i, f in enumerate(ids)}
        return SeriesBatch(np.ascontiguousarray(np.concatenate(parts, axis=0)), np.array([new_row[int(r)] for r in self.row1[keep]], np.int64),
                           np.array([new_row[int(r)] for r in self.row2[keep]], np.int64), z(n1), z(n2),
                           np.ascontiguousarray(self.lm1[lm_keep]), np.ascontiguousarray(self.lm2[lm_keep]),
                           z(np.diff(self.lm_off)[keep]), np.array([remap[int(f)] for f in self.frame_pair[keep]], np.int32), self.name)


def make_series(config: int = 4, *, n_frames: int = 16, plants: int = 1, scale: float = 1.0, seed_offset: int = 0,
                growth=(1.01, 1.05)) -> SeriesBatch:
    """``plants`` plants of the segment population of ``config`` (1 tomato, 2 maize, 3 arabidopsis, 4 = 1), each observed in
    ``n_frames`` frames with cumulative growth (length and point count grow by U(growth) per frame, the curvature drifts,
    every frame is moved rigidly a little further and re-sampled: no identity correspondence) -> plants * (n_frames - 1)
    frame pairs.  Landmarks are sampled per frame pair, as the reference does (segment_matching_ml.py:210-229)."""
    kind = {1: "tomato", 2: "maize", 3: "arabidopsis", 4: "tomato", 5: "tomato"}[config]
    S = max(1, int(round({1: 30, 2: 12, 3: 200, 4: 30, 5: 2000}[config] * scale)))
    voxel = VOXEL[kind]
    shape = "ribbon" if kind == "maize" else "tube"
    frames, row1, row2, n1s, n2s, lm1s, lm2s, fps = [], [], [], [], [], [], [], []
    at = 0
    for p in range(plants):
        rng = np.random.default_rng(1000 * config + 500 + 7919 * (seed_offset + p))
        seg = []
        for _ in range(S):
            if kind == "maize":
                L, size, n = rng.uniform(60, 150), rng.uniform(4, 8), int(rng.uniform(3200, 4800))
            elif config == 5:
                L, size, n = rng.uniform(15, 25), rng.uniform(0.8, 2.0), int(rng.uniform(400, 600))
            else:
                L = rng.uniform(15, 60) if kind == "tomato" else rng.uniform(5, 30)
                size = rng.uniform(0.8, 2.0)
                n = int(rng.uniform(300, 1100)) if kind == "tomato" else int(rng.uniform(150, 1000))
            seg.append(dict(L=L, size=size, n=n, kap=rng.uniform(-0.02, 0.02), Rp=_rot(rng.normal(size=3), rng.uniform(0, np.pi)),
                            tp=rng.uniform(-60, 60, 3), Rm=np.eye(3), tm=np.zeros(3), mirrored=bool(rng.random() < 0.3)))
        prev = None                                          # rows of the previous frame's segments
        for f in range(n_frames):
            rows, arcs = [], []
            for g in seg:
                pts = _surface(rng, g["n"], g["L"], g["kap"], shape, g["size"])
                pts = (pts @ g["Rm"].T + g["tm"]) @ g["Rp"].T + g["tp"]
                frames.append(pts)
                rows.append((at, g["n"]))
                at += g["n"]
                arcs.append((g["L"], g["kap"], g["Rm"].copy(), g["tm"].copy()))
            if prev is not None:
                for (r1, c1), (r2, c2), a1, a2, g in zip(prev[0], rows, prev[1], arcs, seg):
                    l1, l2 = _landmarks(rng, a1[0], a1[1], a2[0], a2[1], voxel, g["mirrored"])
                    lm1s.append((l1 @ a1[2].T + a1[3]) @ g["Rp"].T + g["tp"])
                    lm2s.append((l2 @ a2[2].T + a2[3]) @ g["Rp"].T + g["tp"])
                    row1.append(r1); row2.append(r2); n1s.append(c1); n2s.append(c2)
                    fps.append(p * (n_frames - 1) + f - 1)
            prev = (rows, arcs)
            for g in seg:                                    # grow and move on to the next frame
                gr = rng.uniform(*growth)
                g["L"] *= gr
                g["n"] = int(round(g["n"] * gr))
                g["size"] *= rng.uniform(1.0, 1.01)
                g["kap"] += rng.uniform(-0.002, 0.002)
                dR = _rot(rng.normal(size=3), np.deg2rad(rng.uniform(0, 4)))
                g["Rm"], g["tm"] = dR @ g["Rm"], dR @ g["tm"] + rng.uniform(-0.8, 0.8, 3) * (1 / np.sqrt(3))
    z = lambda n: np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
    k = [len(l) for l in lm1s]
    return SeriesBatch(np.ascontiguousarray(np.concatenate(frames, axis=0)), np.asarray(row1, np.int64), np.asarray(row2, np.int64),
                       z(n1s), z(n2s), np.ascontiguousarray(np.concatenate(lm1s, axis=0)), np.ascontiguousarray(np.concatenate(lm2s, axis=0)),
                       z(k), np.asarray(fps, np.int32), name=f"C{config} series {plants}x{n_frames}")
