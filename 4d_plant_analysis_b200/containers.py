"""Containers of the batched API: CSR batches of matched segment pairs, and growth series that store every frame once.

``SegmentPairBatch`` is what ``local_icp_batch`` / ``match_segments_batch`` take (one or many frame pairs flattened — the inputs of
the reference's loop over matched segment pairs, p2p_matching_in_organ/p2p_matching_seg_icp.py:152-176); ``SeriesBatch`` is the
same for consecutive day pairs of a series (run_registration_pipeline.py:64-104), where frame k is day2 of frame pair k - 1 and
day1 of frame pair k and is therefore stored — and uploaded — once.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Sequence

import numpy as np


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
