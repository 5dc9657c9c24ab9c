#!/bin/bash
# round-2 probe 7: fused write-back; three stream priorities, gather grid, gather ordered behind the sampled aggregate
set -u
OUT=gpurun_out/r02p7
mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_trains_gpu.py tests/test_sampler_train_gpu.py -m gpu -q > "$OUT/pytest_trains.log" 2>&1
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
SGCN_STEP_PRIORITY=1 b k${K}_prio1 --steps $K --warmup 5 --no-cpu --no-also
SGCN_PAD_GRID_MULT=1 b k${K}_grid1 --steps $K --warmup 5 --no-cpu --no-also
SGCN_PAD_GRID_MULT=2 b k${K}_grid2 --steps $K --warmup 5 --no-cpu --no-also
SGCN_GATHER_AFTER_SAMPLED=1 b k${K}_gas --steps $K --warmup 5 --no-cpu --no-also
SGCN_GATHER_AFTER_SAMPLED=1 SGCN_PAD_GRID_MULT=1 b k${K}_gas_grid1 --steps $K --warmup 5 --no-cpu --no-also
b k${K}_nofuse --steps $K --warmup 5 --no-cpu --no-also --no-fuse-write-back
done
FUSE=1 FIRST_TRAIN=4 timeout 120 python tools/timeline.py trains 20 > "$OUT/timeline_fused_prio3.txt" 2>&1; echo "timeline exit $?"; head -45 "$OUT/timeline_fused_prio3.txt"
SGCN_GATHER_AFTER_SAMPLED=1 FUSE=1 FIRST_TRAIN=4 timeout 120 python tools/timeline.py trains 20 > "$OUT/timeline_fused_gas.txt" 2>&1; echo "timeline exit $?"; head -45 "$OUT/timeline_fused_gas.txt"
ls "$OUT"
