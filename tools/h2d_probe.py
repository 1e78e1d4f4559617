"""why is the upload of a small series step slow?  pinned H2D of 10 MB in 8 pieces, alone / next to kernels / with D2H, GB/s per iteration"""
import sys, time, torch
dev = torch.device("cuda", 0)
n = 10 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
ho = torch.empty(n // 3, dtype=torch.uint8).pin_memory()
do = torch.empty(n // 3, dtype=torch.uint8, device=dev)
a = torch.randn(4096, 4096, device=dev)
up, comp = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def run(label, kernels, d2h, gap):
    out = []
    for it in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if kernels:
            with torch.cuda.stream(comp):
                for _ in range(3): (a @ a)
        with torch.cuda.stream(up):
            e0.record(up)
            for k in range(8):
                s = slice(k * n // 8, (k + 1) * n // 8)
                d[s].copy_(h[s], non_blocking=True)
            e1.record(up)
        if d2h:
            with torch.cuda.stream(comp):
                ho.copy_(do, non_blocking=True)
        torch.cuda.synchronize()
        out.append(n / e0.elapsed_time(e1) / 1e6)
        if gap: time.sleep(gap)
    print(f"{label:40s}", " ".join(f"{x:5.1f}" for x in out), "GB/s")
run("copies alone", False, False, 0)
run("copies alone, 2 ms idle between", False, False, 0.002)
run("copies alone, 50 ms idle between", False, False, 0.05)
run("copies next to matmuls", True, False, 0)
run("copies + D2H", False, True, 0)
run("copies + matmuls + D2H", True, True, 0)
