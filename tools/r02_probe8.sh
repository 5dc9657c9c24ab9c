#!/bin/bash
# round-2 probe 8: span rounding (all warps busy), early vs late programmatic trigger under three priorities
set -u
OUT=gpurun_out/r02p8
mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_trains_gpu.py tests/test_aggregate_gpu.py -m gpu -q > "$OUT/pytest_trains.log" 2>&1
echo "pytest exit $?"; tail -3 "$OUT/pytest_trains.log"
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
SGCN_FULL_TRIGGER=0 b k${K}_early --steps $K --warmup 5 --no-cpu --no-also
b k${K}_nofuse --steps $K --warmup 5 --no-cpu --no-also --no-fuse-write-back
SGCN_STEP_PRIORITY=0 b k${K}_nofuse_noprio --steps $K --warmup 5 --no-cpu --no-also --no-fuse-write-back
done
FUSE=1 FIRST_TRAIN=4 timeout 120 python tools/timeline.py trains 20 > "$OUT/timeline_fused_late.txt" 2>&1; echo "timeline exit $?"; sed -n 20,75p "$OUT/timeline_fused_late.txt"
SGCN_FULL_TRIGGER=0 FUSE=1 FIRST_TRAIN=4 timeout 120 python tools/timeline.py trains 20 > "$OUT/timeline_fused_early.txt" 2>&1; echo "timeline exit $?"; sed -n 20,75p "$OUT/timeline_fused_early.txt"
FUSE=0 FIRST_TRAIN=4 timeout 120 python tools/timeline.py trains 20 > "$OUT/timeline_unfused.txt" 2>&1; echo "timeline exit $?"; sed -n 20,60p "$OUT/timeline_unfused.txt"
ls "$OUT"
