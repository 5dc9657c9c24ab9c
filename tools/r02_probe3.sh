#!/bin/bash
# round-2 probe 3: persistent schedule with the gated write-back; conflict-wait diagnostics
set -u
OUT=gpurun_out/r02p3
mkdir -p "$OUT"
for m in probe hiprio late; do timeout 60 python tools/debug_train_wait.py $m > "$OUT/debug_train_wait_$m.txt" 2>&1; head -4 "$OUT/debug_train_wait_$m.txt"; grep " 16," "$OUT/debug_train_wait_$m.txt"; done
timeout 300 python -m pytest tests/test_trains_gpu.py -x -q -k "persistent" > "$OUT/pytest_persistent.log" 2>&1
echo "pytest persistent exit $?"; tail -5 "$OUT/pytest_persistent.log"
b() { name=$1; shift; timeout 300 python bench.py "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"; echo "$name exit $?"; cut -c1-300 "$OUT/$name.json"; tail -3 "$OUT/$name.err"; }
b k20 --steps 20 --warmup 5 --no-cpu --no-also
b k20_persistent --steps 20 --warmup 5 --no-cpu --no-also --persistent
b k2048_persistent --steps 2048 --warmup 5 --no-cpu --no-also --persistent
b reddit_cvd_persistent --workload reddit_cvd --steps 200 --warmup 5 --no-also --no-cpu --persistent
OVERLAP=1 PERSISTENT=1 FIRST_TRAIN=4 timeout 120 python tools/timeline.py trains 20 > "$OUT/timeline_persistent.txt" 2>&1; echo "timeline exit $?"; head -40 "$OUT/timeline_persistent.txt"
timeout 600 python -m pytest tests/test_fullsize_gpu.py -x -q -k "persistent" > "$OUT/pytest_fullsize.log" 2>&1
echo "pytest fullsize exit $?"; tail -5 "$OUT/pytest_fullsize.log"
ls -la "$OUT"
