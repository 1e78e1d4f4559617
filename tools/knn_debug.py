import importlib, os, sys, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden
L = importlib.import_module("4d_plant_analysis_b200")
g = load_golden(sys.argv[1] if len(sys.argv) > 1 else "c1_small")
S = len(g["off1"]) - 1
for s in range(S):
    a1, b1, a2, b2 = g["off1"][s], g["off1"][s + 1], g["off2"][s], g["off2"][s + 1]
    x1, x2 = g["xyz1"][a1:b1], g["xyz2"][a2:b2]
    b = L.batch_from_segments([(x1, x2, x1[:1], x2[:1])])
    r = L.refine_batch(b, g["T21"][a2:b2], g["T12"][a1:b1])
    bad = np.nonzero(r != g["T21r"][a2:b2])[0]
    print("segment", s, "n1", len(x1), "n2", len(x2), "mismatches", len(bad), bad[:5], r[bad[:5]], g["T21r"][a2:b2][bad[:5]])
