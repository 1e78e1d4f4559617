"""Wiring into an unmodified checkout of the reference (SURVEY.md §8b "import-order trap").

    import local_icp_b200; local_icp_b200.install()          # after the reference modules are importable
    python -m local_icp_b200.run_reference /path/to/4d_plant_analysis/run_registration_pipeline.py --type tomato --method local_icp

``install()`` replaces, in the reference's already-imported (or importable) modules, the five hot
functions by the GPU versions.  The reference resolves them as module globals at call time
(p2p_matching_seg_icp.py:173,175,416,446), so patching any time before the call is enough and no
reference file is edited.
"""
from __future__ import annotations

import importlib
import sys

from . import p2p_matching_seg_icp as _seg
from . import registration_icp as _reg

_REG_NAMES = ("get_best_fit_transform", "get_NN_in_target", "my_icp")
_SEG_NAMES = ("get_triangle_angles", "get_p2p_mapping_from_segments_using_icp")


def _module(name):
    if name in sys.modules:
        return sys.modules[name]
    try:
        return importlib.import_module(name)
    except Exception:        # ImportError, or an OSError / AttributeError from an eager open3d or pyFM import in the reference
        return None


def install(verbose: bool = False):
    """Patch the reference's modules in place.  Returns the list of (module, name) patched."""
    done = []
    reg = _module("registration.registration_icp")
    if reg is not None:
        for n in _REG_NAMES:
            setattr(reg, n, getattr(_reg, n))
            done.append((reg.__name__, n))
    seg = _module("p2p_matching_in_organ.p2p_matching_seg_icp")
    if seg is not None:
        for n in _SEG_NAMES:
            setattr(seg, n, getattr(_seg, n))
            done.append((seg.__name__, n))
        seg.my_icp = _reg.my_icp                       # imported by name at p2p_matching_seg_icp.py:24
        done.append((seg.__name__, "my_icp"))
    ev = _module("p2p_matching_in_organ.p2p_matching_evaluation")
    if ev is not None:
        from . import p2p_matching_evaluation as _ev
        ev.evaluate_p2p_matching = _ev.evaluate_p2p_matching
        done.append((ev.__name__, "evaluate_p2p_matching"))
        if seg is not None and hasattr(seg, "evaluate_p2p_matching"):   # imported by name into the driver module
            seg.evaluate_p2p_matching = _ev.evaluate_p2p_matching
            done.append((seg.__name__, "evaluate_p2p_matching"))
    fm = _module("p2p_matching_in_organ.p2p_matching_fm")
    if fm is not None:
        from . import p2p_matching_fm as _fm
        fm.double_direction_refinement = _fm.double_direction_refinement
        done.append((fm.__name__, "double_direction_refinement"))
        if seg is not None and hasattr(seg, "double_direction_refinement"):   # imported by name into the driver module
            seg.double_direction_refinement = _fm.double_direction_refinement
            done.append((seg.__name__, "double_direction_refinement"))
    ip = sys.modules.get("interpolation.interpolation_pcd")     # only if the caller imported it (it needs open3d, imageio)
    if ip is not None:
        from . import interpolation_pcd as _ip
        ip.interpolation_segment_level = _ip.interpolation_segment_level
        done.append((ip.__name__, "interpolation_segment_level"))
    if verbose:
        for m, n in done:
            print(f"local_icp_b200: patched {m}.{n}")
    return done
