import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
L = importlib.import_module("4d_plant_analysis_b200")
b = L.make_config(int(sys.argv[1]) if len(sys.argv) > 1 else 2, frame_pairs=int(sys.argv[2]) if len(sys.argv) > 2 else 16)
r = L.match_segments_batch(b)
print("ok", r.score)
