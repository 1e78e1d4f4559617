"""Batched entry point (not in the reference): all matched segment pairs of one or
many frame pairs in ONE pass over the GPU pipeline, both directions.  Everything the drop-in
functions do is a batch of one through the same code.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from .containers import SegmentPairBatch, batch_from_segments


@dataclass
class LocalICPResult:
    """Dense maps and per-landmark ICP results of a batch.

    T21: (sum N2,) int64, T12: (sum N1,) int64 (concatenated in batch order, local indices);
    transforms: dict with R21/t21/iters21/status21 and the 12 twins, one row per landmark
    (after the fill step), or None when not requested."""

    T21: Optional[np.ndarray]
    T12: Optional[np.ndarray]
    off1: np.ndarray
    off2: np.ndarray
    lm_off: np.ndarray
    transforms: Optional[dict] = None
    counters: Optional[dict] = None

    def maps(self, s: int) -> Tuple[np.ndarray, np.ndarray]:
        """(T21, T12) of segment pair s, as the reference's pair of calls returns them."""
        return (self.T21[self.off2[s]:self.off2[s + 1]] if self.T21 is not None else None,
                self.T12[self.off1[s]:self.off1[s + 1]] if self.T12 is not None else None)

    def as_list(self) -> List[Tuple[np.ndarray, np.ndarray]]:
        return [self.maps(s) for s in range(len(self.off1) - 1)]


def _as_batch(segments) -> SegmentPairBatch:
    if isinstance(segments, SegmentPairBatch):
        return segments
    return batch_from_segments([(np.asarray(a, np.float64), np.asarray(b, np.float64), np.asarray(c, np.float64),
                                 np.asarray(d, np.float64)) for a, b, c, d in segments])


def local_icp_batch(segments, *, directions: int = _native.DIR_BOTH, details: bool = False,
                    device: Optional[int] = None, validate: bool = True, ctx: Optional[_native.Context] = None,
                    ) -> LocalICPResult:
    """segments: a SegmentPairBatch or a list of (xyz1, xyz2, lm1, lm2).  Host numpy in, host numpy out."""
    b = _as_batch(segments)
    ctx = ctx or _native.default_context(device)
    with ctx.lock:
        ctx.set_input_validation(validate)   # NaN / infinity -> ValueError, checked by the library while it stages the arrays
        try:
            ctx.plan(b.off1, b.off2, b.lm_off, directions)
            T21, T12 = ctx.run_host(b.xyz1, b.xyz2, b.lm1, b.lm2)
        finally:
            ctx.set_input_validation(False)
        res = LocalICPResult(T21, T12, b.off1, b.off2, b.lm_off)
        if details:
            res.transforms = ctx.transforms(filled=True)
            res.transforms["raw"] = ctx.transforms(filled=False)
            res.counters = ctx.counters()
    return res


@dataclass
class SegmentMatchResult:
    """What the reference's per-segment body (p2p_matching_seg_icp.py:152-196) accumulates for one frame pair."""

    T21s: List[np.ndarray]            # per matched segment pair, after the refinement if requested
    T12s: List[np.ndarray]
    scores: np.ndarray                # (S, 3): continuity, landmark precision, cycle consistency  (score.csv)
    score_weights: np.ndarray         # (S,): vertices of mesh1 (:189)
    segments: SegmentPairBatch

    @property
    def score(self) -> np.ndarray:    # what get_p2p_mapping_organ_collection returns (:244)
        return np.average(self.scores, axis=0, weights=self.score_weights)


def match_segments_batch(segments, *, apply_double_direction_refinement: bool = True, voxel_size: float = 1.0,
                         sample_proportion: float = 0.5, ctx: Optional[_native.Context] = None) -> SegmentMatchResult:
    """The whole per-segment body of get_p2p_mapping_organ_collection for all matched segment pairs at once, on the GPU:
    T21 / T12 (p2p_matching_seg_icp.py:173-176), double_direction_refinement (:181-182), evaluate_p2p_matching with the
    driver's arguments (:184-185: default voxel_size = 1.0, sample_proportion = 0.5) and the score weights (:189)."""
    b = _as_batch(segments)
    ctx = ctx or _native.default_context()
    with ctx.lock:
        ctx.set_input_validation(True)
        try:
            ctx.plan(b.off1, b.off2, b.lm_off)
            T21, T12, scores = ctx.match(b.xyz1, b.xyz2, b.lm1, b.lm2, apply_double_direction_refinement, voxel_size, sample_proportion)
        finally:
            ctx.set_input_validation(False)
    S = b.n_segments
    return SegmentMatchResult([T21[b.off2[s]:b.off2[s + 1]] for s in range(S)], [T12[b.off1[s]:b.off1[s + 1]] for s in range(S)],
                              scores, np.diff(b.off1).astype(np.float64), b)


def save_segment_results(result: SegmentMatchResult, directory: str, day1: str, day2: str) -> None:
    """The on-disk format the reference's driver writes (p2p_matching_seg_icp.py:210-242) and its consumers read
    (interpolation/interpolation_pcd.py:15-55, visualize_p2p.py:34-74): per matched segment pair i
    ``{day1}_{i}.csv`` / ``{day2}_{i}.csv`` (vertices, np.savetxt default format), ``T21_{i}.csv`` / ``T12_{i}.csv``
    (one float per line) and ``score.csv``."""
    import os
    os.makedirs(directory, exist_ok=True)
    b = result.segments
    for i in range(b.n_segments):
        x1, x2, _, _ = b.segment(i)
        np.savetxt(os.path.join(directory, f"{day1}_{i}.csv"), x1)
        np.savetxt(os.path.join(directory, f"{day2}_{i}.csv"), x2)
        np.savetxt(os.path.join(directory, f"T21_{i}.csv"), result.T21s[i])
        np.savetxt(os.path.join(directory, f"T12_{i}.csv"), result.T12s[i])
    np.savetxt(os.path.join(directory, "score.csv"), np.vstack(result.scores))
