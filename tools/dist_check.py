"""torchrun --nproc-per-node N tools/dist_check.py : shards a growth series (by frame pair, boundary frames over NVLink) and a
one-frame-pair batch (by cost) over N GPUs through ShardedLocalICP (NCCL gather of maps + score counts) and checks, on every
rank, that the gathered results equal the whole batch run on this rank's GPU alone."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
L = importlib.import_module("4d_plant_analysis_b200")
D = importlib.import_module("4d_plant_analysis_b200.distributed")
syn = importlib.import_module("4d_plant_analysis_b200.synthetic")
for name, data in (("series", syn.make_series(4, n_frames=2 * dist.get_world_size() + 1, scale=0.4)), ("c5", syn.make_config(5, scale=0.05))):
    b = data.as_batch() if hasattr(data, "frames") else data
    ref = L.local_icp_batch(b, ctx=L.Context(local))
    sc = L.evaluate_batch(b, ref.T21, ref.T12, voxel_size=1.0, sample_proportion=0.5, ctx=L.Context(local))
    sh = D.ShardedLocalICP(data, device=local)
    for step in (sh.step, sh.step_from_host, sh.step_from_host):
        T21, T12, cnt = step()
        sh.stream.synchronize()
        assert np.array_equal(T21.cpu().numpy(), ref.T21) and np.array_equal(T12.cpu().numpy(), ref.T12), f"{name}: gathered maps differ"
        assert np.allclose(D.scores_from_counts(cnt.cpu().numpy()), sc, rtol=0, atol=1e-15), f"{name}: gathered scores differ"
    if name == "series" and dist.get_world_size() > 1:
        assert (sh.boundary_rows > 0) == (dist.get_rank() + 1 < dist.get_world_size()), "boundary frame must come from the next rank"
    print(f"rank {dist.get_rank()} {name}: {len(sh.mine)} of {data.n_segments} segment pairs, boundary rows {sh.boundary_rows}")
print("dist_check ok")
dist.barrier()
dist.destroy_process_group()
