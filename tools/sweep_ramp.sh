# sizes of the sub-batches of the host pipeline: cut k where the running vertex count passes (k/chunks)^ramp of the total
for rep in 1 2; do for ch in ${CHUNKS:-8 6}; do for ramp in ${RAMPS:-1.0 1.3 1.6 2.0}; do
LIP_PIPELINE_CHUNKS=$ch LIP_PIPELINE_RAMP=$ramp python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sharded --no-parity --no-configs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunks $ch ramp $ramp: e2e', round(d['e2e']['ms_per_step'],3), {k: round(v['ms_per_step'],2) for k,v in d['e2e']['variants'].items() if k != 'pageable_new_shape_validated'})"
done; done; done
