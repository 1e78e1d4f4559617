# sub-batches of the host pipeline on every bench config (e2e ms per step)
for rep in 1 2; do for ch in 8 4 2; do
LIP_PIPELINE_CHUNKS=$ch python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sharded --no-parity > gpurun_out/ch.json 2>gpurun_out/ch.err; python - $ch <<PY
import json, sys
d=json.loads(open("gpurun_out/ch.json").read().strip().splitlines()[-1])
print("chunks", sys.argv[1], "headline e2e", round(d["e2e"]["ms_per_step"],3), [(c["config"], round(c["ms_per_step"],3), round(c["e2e_ms_per_step"],3)) for c in d["configs"]])
PY
done; done
