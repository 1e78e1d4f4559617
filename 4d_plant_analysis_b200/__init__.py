"""B200-native local-ICP point-to-point matching (the hot path of romi/4d_plant_analysis).

The directory name starts with a digit, so import it with
``importlib.import_module("4d_plant_analysis_b200")`` or through the root-level alias
``import local_icp_b200``.
"""
