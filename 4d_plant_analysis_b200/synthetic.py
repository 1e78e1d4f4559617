"""Seeded synthetic plant segments for the five BASELINE.json configs.

Nothing here is on the product path: it only manufactures inputs with the shape of what the
reference's coarse matching hands to the hot path — per matched segment pair two vertex sets
(``mesh1.vertices``, ``mesh2.vertices``) and two equally long landmark polylines (``lm1``, ``lm2``),
see ``p2p_matching_in_organ/p2p_matching_seg_icp.py:152-176`` and the landmark thinning in
``segment_matching/segment_matching_ml.py:210-229`` of the reference.

Everything is generated in float64 with ``np.random.default_rng(1000 * config + frame_index)``.
"""
from __future__ import annotations

import numpy as np

from .containers import SegmentPairBatch, SeriesBatch, batch_from_segments, concat_batches  # noqa: F401  (re-exported)

VOXEL = {"tomato": 0.8, "maize": 0.8, "arabidopsis": 0.7}


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
