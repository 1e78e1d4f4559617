timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for rep in 1 2; do for gr in 0 1; do
LIP_GRAPHS=$gr python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sharded --no-parity > gpurun_out/gr.json 2>gpurun_out/gr.err; python - $gr <<PY
import json, sys
d=json.loads(open("gpurun_out/gr.json").read().strip().splitlines()[-1])
print("graphs", sys.argv[1], "headline", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["ms_per_step"],3), {k: round(v["ms_per_step"],2) for k,v in d["e2e"]["variants"].items()}, [(c["config"], round(c["ms_per_step"],3), round(c["e2e_ms_per_step"],3)) for c in d["configs"]])
PY
done; done
for gr in 0 1; do LIP_GRAPHS=$gr python tools/trace_series.py 4 16 1 2>&1 | grep step | tail -2; done
