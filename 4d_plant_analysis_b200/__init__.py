"""B200-native local-ICP point-to-point matching (the hot path of romi/4d_plant_analysis).

The directory name starts with a digit, so import it with
``importlib.import_module("4d_plant_analysis_b200")`` or through the root-level alias
``import local_icp_b200``.

Reference-shaped API (same names and signatures as the reference's functions):
    my_icp, get_best_fit_transform, get_NN_in_target                 (registration/registration_icp.py)
    get_triangle_angles, get_p2p_mapping_from_segments_using_icp     (p2p_matching_in_organ/p2p_matching_seg_icp.py)
    evaluate_p2p_matching                                            (p2p_matching_in_organ/p2p_matching_evaluation.py)
    double_direction_refinement                                      (p2p_matching_in_organ/p2p_matching_fm.py)
    interpolation_segment_level                                      (interpolation/interpolation_pcd.py)
    label_points_by_skeleton, skeleton_landmarks                     (segment_matching/segment_matching_utils.py, _ml.py)
Batched API: local_icp_batch, refine_batch, evaluate_batch, match_segments_batch (the whole per-segment body of the
reference's driver), save_segment_results (its on-disk format).  Wiring into a reference checkout: install().
"""
from ._native import DIR_BOTH, DIR_T12, DIR_T21, Context, NativeError, build, default_context, load  # noqa: F401
from .batch import (LocalICPResult, SegmentMatchResult, local_icp_batch, match_segments_batch,  # noqa: F401
                    save_segment_results)
from .install import install  # noqa: F401
from .interpolation_pcd import interpolate_batch, interpolate_saved_result, interpolation_segment_level  # noqa: F401
from .p2p_matching_evaluation import evaluate_batch, evaluate_p2p_matching  # noqa: F401
from .p2p_matching_fm import double_direction_refinement, refine_batch  # noqa: F401
from .p2p_matching_seg_icp import get_p2p_mapping_from_segments_using_icp, get_triangle_angles  # noqa: F401
from .registration_icp import get_best_fit_transform, get_NN_in_target, my_icp  # noqa: F401
from .segment_matching_utils import label_points_by_skeleton, skeleton_landmarks, split_by_label  # noqa: F401
from .containers import SegmentPairBatch, SeriesBatch, batch_from_segments, concat_batches  # noqa: F401
from .synthetic import make_config, make_series  # noqa: F401
from .distributed import ShardedLocalICP, gather_results, local_icp_distributed, make_shard_plan, shard_segments  # noqa: F401
