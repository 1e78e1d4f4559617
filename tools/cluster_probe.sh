export LIP_DEVICE_CHUNKS=1
for cfg in 2 3; do for fp in 1 4; do for cc in 4 2 1 0; do
  echo "== config $cfg x $fp frame pairs, cluster class $cc: $(LIP_CLUSTER_CLASS=$cc LIP_CLUSTER_COOP_PARTS=100000 python tools/profile_config.py $cfg $fp 3 2>&1 | tail -1 | cut -c1-70)"
done; done; done
for cc in 4 1; do
  LIP_CLUSTER_CLASS=$cc LIP_CLUSTER_COOP_PARTS=100000 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/cfg2_cc$cc.csv python tools/profile_config.py 2 4 1 > /dev/null 2>&1
  python - $cc <<'PY'
import csv, sys, collections
rows = [r for r in csv.reader(open("gpurun_out/cfg2_cc%s.csv" % sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
acc = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0]; acc.setdefault(name, []).append(float(r[-1].replace(",", "")))
print("C2x4 class", sys.argv[1], {k[-24:]: round(v[-1]/1e3, 1) for k, v in acc.items()})
PY
done
