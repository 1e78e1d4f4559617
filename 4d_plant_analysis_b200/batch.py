"""Batched entry point (not in the reference): all matched segment pairs of one or
many frame pairs in ONE pass over the GPU pipeline, both directions.  Everything the drop-in
functions do is a batch of one through the same code.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _native
from .synthetic import SegmentPairBatch, batch_from_segments


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
    if validate:
        for a in (b.xyz1, b.xyz2, b.lm1, b.lm2):
            if not np.isfinite(a).all():
                raise ValueError("Input contains NaN or infinity.")
    ctx = ctx or _native.default_context(device)
    ctx.plan(b.off1, b.off2, b.lm_off, directions)
    T21, T12 = ctx.run_host(b.xyz1, b.xyz2, b.lm1, b.lm2)
    res = LocalICPResult(T21, T12, b.off1, b.off2, b.lm_off)
    if details:
        res.transforms = ctx.transforms(filled=True)
        res.transforms["raw"] = ctx.transforms(filled=False)
        res.counters = ctx.counters()
    return res
