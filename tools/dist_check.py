"""torchrun --nproc-per-node N tools/dist_check.py : shards a batch over N GPUs (local_icp_distributed, NCCL gather) and
checks the gathered maps against a single-GPU run of the whole batch on every rank."""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = importlib.import_module("4d_plant_analysis_b200")
D = importlib.import_module("4d_plant_analysis_b200.distributed")
for cfg, kw in ((4, dict(scale=0.3)), (5, dict(scale=0.05))):          # frame-pair sharding (C4) and segment-pair LPT sharding (C5)
    b = L.make_config(cfg, **kw)
    T21, T12, mine = D.local_icp_distributed(b, gather=True)
    ref = L.local_icp_batch(b)
    assert np.array_equal(T21, ref.T21) and np.array_equal(T12, ref.T12), f"config {cfg}: gathered maps differ"
    print(f"rank {dist.get_rank()}/{dist.get_world_size()} config {cfg}: {len(mine)} of {b.n_segments} segment pairs local, gather ok", flush=True)
dist.barrier()
dist.destroy_process_group()
