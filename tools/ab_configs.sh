# tools/ab.sh on every bench config: AB_LIBS="name=path ..." bash tools/ab_configs.sh
AB_ARGS="--no-sharded" bash tools/ab.sh $AB_LIBS
for v in $AB_LIBS; do python - ${v%%=*} <<PY
import json, sys
d=json.loads(open("gpurun_out/ab_%s_2.json" % sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d["parity"]["map_mismatch"], d["parity"]["iter_identical_frac"], [(c["config"], round(c["ms_per_step"],3), round(c["e2e_ms_per_step"],3), c["stage_ms"]["icp_group_kernel"], c["stage_ms"]["icp_big_kernel"]) for c in d["configs"]])
PY
done
