for c in "1 1" "2 1" "3 1" "1 4" "2 4" "2 64"; do timeout 300 python tools/graph_probe.py $c 2>&1 | tail -3; done
