timeout 300 python -m pytest tests -m gpu -x -q -k "refine or match_segments" 2>&1 | tail -3
for lib in cur refw; do
  LIP_LIBRARY=build/$lib.so ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ref_$lib.csv python tools/refine_step.py 2 64 > /dev/null 2>&1
  python - $lib <<'PY'
import csv, sys
rows = [r for r in csv.reader(open("gpurun_out/ref_%s.csv" % sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
print(sys.argv[1], [(r[4].split("(")[0][-40:], round(float(r[-1].replace(",", ""))/1e6, 3)) for r in rows if "refine" in r[4] or "self_knn" in r[4]])
PY
done
