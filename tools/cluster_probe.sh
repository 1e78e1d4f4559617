AB_ARGS="--no-sharded --no-parity" bash tools/ab.sh cur=build/cur.so boxb32=build/sub4.so c384b64=build/c384b64.so c384b32=build/c384b32.so
for v in cur boxb32 c384b64 c384b32; do python - $v <<PY
import json, sys
d=json.loads(open("gpurun_out/ab_%s_2.json" % sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], [(c["config"], round(c["ms_per_step"],3), round(c["e2e_ms_per_step"],3), c["stage_ms"]["final_map_kernel"]) for c in d["configs"]], {k: round(x["ms_per_step"],2) for k,x in d["e2e"]["variants"].items()})
PY
done
