#!/bin/bash
# round-2 probe 5: write-back fused into the tail of the full-neighbour mean; L2 eviction policies
set -u
OUT=gpurun_out/r02p5
mkdir -p "$OUT"
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_fullsize_gpu.py > "$OUT/pytest_gpu.log" 2>&1
echo "pytest gpu exit $?"; tail -15 "$OUT/pytest_gpu.log"
b() { name=$1; shift; timeout 300 python bench.py "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"; echo "$name exit $?"; python - "$OUT/$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ms/step %.5f  e2e %.5f  full_mean us %.2f  frac %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["us_per_launch"], d["roofline"]["frac"]))
except Exception as e: print("   parse failed", e)
PY
tail -3 "$OUT/$name.err"; }
b k20 --steps 20 --warmup 5 --no-cpu --no-also
b k20_nofuse --steps 20 --warmup 5 --no-cpu --no-also --no-fuse-write-back
b k2048 --steps 2048 --warmup 5 --no-cpu --no-also
b k2048_nofuse --steps 2048 --warmup 5 --no-cpu --no-also --no-fuse-write-back
SGCN_HIST_L2=100 b k2048_h100 --steps 2048 --warmup 5 --no-cpu --no-also
SGCN_HIST_L2=75 b k2048_h75 --steps 2048 --warmup 5 --no-cpu --no-also
SGCN_HIST_L2=50 b k2048_h50 --steps 2048 --warmup 5 --no-cpu --no-also
SGCN_STREAM_L2=100 b k2048_s100 --steps 2048 --warmup 5 --no-cpu --no-also
SGCN_HIST_L2=100 SGCN_STREAM_L2=100 b k2048_h100_s100 --steps 2048 --warmup 5 --no-cpu --no-also
SGCN_HIST_L2=75 SGCN_STREAM_L2=100 b k2048_h75_s100 --steps 2048 --warmup 5 --no-cpu --no-also
SGCN_HIST_L2=100 SGCN_STREAM_L2=100 b k20_h100_s100 --steps 20 --warmup 5 --no-cpu --no-also
FUSE=1 FIRST_TRAIN=4 timeout 120 python tools/timeline.py trains 20 > "$OUT/timeline_fused.txt" 2>&1; echo "timeline exit $?"; head -60 "$OUT/timeline_fused.txt"
ls "$OUT"
