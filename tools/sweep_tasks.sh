# group size of icp_group_kernel on every bench config: LIP_ICP_TASKS_PER_WARP x LIP_ICP_CLASS_SHIFT
for tpw in 1 2 3 4; do for csh in 4 3; do
LIP_ICP_TASKS_PER_WARP=$tpw LIP_ICP_CLASS_SHIFT=$csh python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sharded --no-parity > gpurun_out/tpw.json 2>gpurun_out/tpw.err; python - $tpw $csh <<PY
import json, sys
d=json.loads(open("gpurun_out/tpw.json").read().strip().splitlines()[-1])
print("tasks/warp", sys.argv[1], "class shift", sys.argv[2], "headline", round(d["ms_per_step"],3), round(d["stage_ms"]["icp_group_kernel"],3), [(c["config"], round(c["ms_per_step"],3), c["stage_ms"]["icp_group_kernel"]) for c in d["configs"]])
PY
done; done
