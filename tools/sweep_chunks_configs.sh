# sub-batches of the host pipeline on every bench config (e2e ms per step): "chunks:ramp" pairs in $SHAPES
for rep in 1 2; do for sh in ${SHAPES:-8:1.0 8:1.6 7:2.0}; do
LIP_PIPELINE_CHUNKS=${sh%%:*} LIP_PIPELINE_RAMP=${sh##*:} python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sharded --no-parity > gpurun_out/ch.json 2>gpurun_out/ch.err; python - $sh <<PY
import json, sys
d=json.loads(open("gpurun_out/ch.json").read().strip().splitlines()[-1])
print("shape", sys.argv[1], "headline e2e", round(d["e2e"]["ms_per_step"],3), [(c["config"], round(c["e2e_ms_per_step"],3)) for c in d["configs"]], round(d["body_e2e"]["ms_per_step"],1))
PY
done; done
