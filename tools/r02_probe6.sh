#!/bin/bash
# round-2 probe 6: fused write-back with the late programmatic trigger and prioritised side streams
set -u
OUT=gpurun_out/r02p6
mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_trains_gpu.py tests/test_sampler_train_gpu.py -m gpu -q -x > "$OUT/pytest_trains.log" 2>&1
echo "pytest trains exit $?"; tail -5 "$OUT/pytest_trains.log"
b() { name=$1; shift; timeout 300 python bench.py "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"; echo "$name exit $?"; python - "$OUT/$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ms/step %.5f  e2e %.5f  full_mean us %.2f  frac %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["us_per_launch"], d["roofline"]["frac"]))
except Exception as e: print("   parse failed", e)
PY
tail -3 "$OUT/$name.err"; }
for K in 20 2048; do
b k${K} --steps $K --warmup 5 --no-cpu --no-also
SGCN_STEP_PRIORITY=0 b k${K}_noprio --steps $K --warmup 5 --no-cpu --no-also
SGCN_FULL_TRIGGER=0 b k${K}_early --steps $K --warmup 5 --no-cpu --no-also
b k${K}_nofuse --steps $K --warmup 5 --no-cpu --no-also --no-fuse-write-back
SGCN_STEP_PRIORITY=0 b k${K}_nofuse_noprio --steps $K --warmup 5 --no-cpu --no-also --no-fuse-write-back
SGCN_FULL_TRIGGER=2 b k${K}_nofuse_late --steps $K --warmup 5 --no-cpu --no-also --no-fuse-write-back
done
FUSE=1 FIRST_TRAIN=4 timeout 120 python tools/timeline.py trains 20 > "$OUT/timeline_fused_late_prio.txt" 2>&1; echo "timeline exit $?"; head -45 "$OUT/timeline_fused_late_prio.txt"
SGCN_STEP_PRIORITY=0 FUSE=1 FIRST_TRAIN=4 timeout 120 python tools/timeline.py trains 20 > "$OUT/timeline_fused_late_noprio.txt" 2>&1; echo "timeline exit $?"; head -45 "$OUT/timeline_fused_late_noprio.txt"
ls "$OUT"
