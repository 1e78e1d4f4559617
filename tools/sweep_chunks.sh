for pc in 4 6 8 12; do for dc in 2; do
LIP_PIPELINE_CHUNKS=$pc LIP_DEVICE_CHUNKS=$dc python bench.py --steps 10 --warmup 3 --no-sharded --no-cpu-baseline --no-configs --no-parity > gpurun_out/sw_${pc}_${dc}.json 2>/dev/null
python - $pc $dc <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/sw_{sys.argv[1]}_{sys.argv[2]}.json').read().strip().splitlines()[-1])
print('pipeline chunks', sys.argv[1], 'device chunks', sys.argv[2], 'dev %.3f ms e2e %.3f ms pageable %.3f ms' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['variants']['pageable_numpy']['ms_per_step']))
PY
done; done
for dc in 1 3 4; do
LIP_DEVICE_CHUNKS=$dc python bench.py --steps 10 --warmup 3 --no-sharded --no-cpu-baseline --no-configs --no-parity > gpurun_out/sw_d${dc}.json 2>/dev/null
python - $dc <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/sw_d{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('device chunks', sys.argv[1], 'dev %.3f ms e2e %.3f ms' % (d['ms_per_step'], d['e2e']['ms_per_step']))
PY
done
