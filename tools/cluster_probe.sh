timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export LIP_DEVICE_CHUNKS=1
for cfg in 3 1 2 5; do for cs in 1 8; do
  echo "== config $cfg x 1 frame pair, cluster size $cs: $(LIP_CLUSTER_SIZE=$cs python tools/profile_config.py $cfg 1 3 2>&1 | tail -1 | cut -c1-100)"
done; done
for fp in 2 4; do for cs in 1 8; do
  echo "== config 3 x $fp frame pair, cluster size $cs: $(LIP_CLUSTER_SIZE=$cs python tools/profile_config.py 3 $fp 3 2>&1 | tail -1 | cut -c1-100)"
done; done
unset LIP_DEVICE_CHUNKS
for rep in 1 2; do for cs in 1 8; do LIP_CLUSTER_SIZE=$cs python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sharded --no-parity > gpurun_out/cl_$cs.json 2>gpurun_out/cl_$cs.err; python - $cs <<PY
import json, sys
cs=int(sys.argv[1])
d=json.loads(open("gpurun_out/cl_%d.json"%cs).read().strip().splitlines()[-1])
print(cs, "headline", round(d["ms_per_step"],3), round(d["e2e"]["ms_per_step"],3), d["stage_ms"]["icp_big_kernel"], [(c["config"], round(c["ms_per_step"],3), round(c["e2e_ms_per_step"],3)) for c in d["configs"]])
PY
done; done
