"""torchrun --nproc-per-node N tools/pcie_probe.py : aggregate host<->device copy bandwidth with N ranks copying at once (pinned
buffers of the size of one bench step): H2D alone, D2H alone, both directions together.  Names the limiter of the end-to-end
scaling of bench.py at N GPUs."""
import os, time
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
MB_IN, MB_OUT, REP = 102, 32, 20
h_in = torch.empty(MB_IN << 20, dtype=torch.uint8).pin_memory(); d_in = torch.empty_like(h_in, device=dev)
h_out = torch.empty(MB_OUT << 20, dtype=torch.uint8).pin_memory(); d_out = torch.empty_like(h_out, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
def barrier():
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
def run(h2d, d2h):
    barrier()
    t = time.perf_counter()
    for _ in range(REP):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t], device=dev)
    if world > 1: dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return float(dt)
for _ in range(2): run(True, True)
a, b, c = run(True, False), run(False, True), run(True, True)
if local == 0:
    gb = lambda mb, t: world * REP * mb / 1024 / t
    print(f"N={world}: H2D alone {gb(MB_IN, a):.1f} GB/s aggregate ({gb(MB_IN, a) / world:.1f} per GPU) | D2H alone {gb(MB_OUT, b):.1f} ({gb(MB_OUT, b) / world:.1f} per GPU) | "
          f"both: {gb(MB_IN + MB_OUT, c):.1f} aggregate = {1e3 * c / REP:.2f} ms per (102 MB in + 32 MB out) step")
if world > 1:
    dist.barrier(); dist.destroy_process_group()
