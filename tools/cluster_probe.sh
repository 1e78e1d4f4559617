timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export LIP_DEVICE_CHUNKS=1
for lib in cur wide4; do for c in "1 1" "2 1" "3 1" "2 4" "3 4"; do
  echo "== $lib config $c: $(LIP_LIBRARY=build/$lib.so python tools/profile_config.py $c 3 2>&1 | tail -1 | cut -c1-16)"
done; done
unset LIP_DEVICE_CHUNKS
AB_ARGS="--no-sharded --no-parity --no-configs" bash tools/ab.sh cur=build/cur.so wide4=build/wide4.so
