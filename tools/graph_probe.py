"""does replaying one planned step as a CUDA graph pay?  device-resident step of `config x frame_pairs`, plain launches vs a graph
captured around the same call (torch.cuda.graph on the caller's stream; the library's side streams fork from / join it through events)
   python tools/graph_probe.py 1 1"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
native = importlib.import_module("4d_plant_analysis_b200._native")
syn = importlib.import_module("4d_plant_analysis_b200.synthetic")
cfg, fp = int(sys.argv[1]), int(sys.argv[2])
data = syn.make_config(cfg, frame_pairs=fp)
ctx = native.Context(0)
ctx.plan(data.off1, data.off2, data.lm_off)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(a).to(dev) for a in (data.xyz1, data.xyz2, data.lm1, data.lm2)]
T21 = torch.empty(int(data.off2[-1]), dtype=torch.int64, device=dev); T12 = torch.empty(int(data.off1[-1]), dtype=torch.int64, device=dev)
st = torch.cuda.Stream(dev)
def step():
    ctx.run_device(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), T21.data_ptr(), T12.data_ptr(), st.cuda_stream)
def timed(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(n): fn()
        e1.record(st)
    t_cpu = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, t_cpu * 1e3
ms, cpu = timed(step)
ref21, ref12 = T21.clone(), T12.clone()
print(f"config {cfg} x {fp}: launches  {ms*1e3:8.1f} us/step  (host enqueue {cpu*1e3:6.1f} us)")
g = torch.cuda.CUDAGraph()
T21.zero_(); T12.zero_()
torch.cuda.synchronize()
with torch.cuda.graph(g, stream=st, capture_error_mode="thread_local"):
    step()
def replay():
    g.replay()
with torch.cuda.stream(st):
    ms, cpu = timed(replay)
print(f"config {cfg} x {fp}: graph     {ms*1e3:8.1f} us/step  (host enqueue {cpu*1e3:6.1f} us)  identical maps: {bool((T21 == ref21).all() and (T12 == ref12).all())}")
