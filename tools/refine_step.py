"""One refinement + evaluation + interpolation of a synthetic batch (for ncu launch lists): python tools/refine_step.py [config] [frame_pairs]"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module("4d_plant_analysis_b200")
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
fp = int(sys.argv[2]) if len(sys.argv) > 2 else 64
b = L.make_config(cfg, frame_pairs=fp)
res = L.local_icp_batch(b, validate=False)
T21r = L.refine_batch(b, res.T21, res.T12)
sc = L.evaluate_batch(b, T21r, res.T12, voxel_size=0.8, sample_proportion=0.5)
xi = L.interpolate_batch(b, T21r, res.T12, 0.5)
print("ok", (T21r != res.T21).mean(), sc.mean(0), xi.shape)
