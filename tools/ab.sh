#!/bin/bash
# A/B of library variants on ONE box: tools/ab.sh name=path.so [name=path.so ...]  (two repeats each, bench.py defaults)
# prints per variant: device ms/step, e2e ms/step and the per-kernel stage times
mkdir -p gpurun_out
for rep in 1 2; do
  for v in "$@"; do
    name=${v%%=*}; lib=${v#*=}
    LIP_LIBRARY=$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline $AB_ARGS > gpurun_out/ab_${name}_$rep.json 2> gpurun_out/ab_${name}_$rep.err
    python - "$name" "$rep" gpurun_out/ab_${name}_$rep.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
    st = d.get("stage_ms") or {}
    print(sys.argv[1], sys.argv[2], "dev %.3f ms  e2e %.3f ms " % (d["ms_per_step"], d["e2e"]["ms_per_step"]),
          {k: round(v, 3) for k, v in st.items()} if isinstance(st, dict) else st)
except Exception as e:
    print(sys.argv[1], sys.argv[2], "FAILED", e)
PY
  done
done
