#!/bin/bash
# round-2 probe 10: tcgen05 gathered dense layer; late programmatic trigger of the write-back kernels
set -u
OUT=gpurun_out/r02p10
mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > "$OUT/pytest_gemm.log" 2>&1
echo "pytest gemm exit $?"; tail -25 "$OUT/pytest_gemm.log" | cut -c1-220
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
SGCN_WB_TRIGGER=0 b k${K}_early --steps $K --warmup 5 --no-cpu --no-also
done
FUSE=0 FIRST_TRAIN=4 timeout 120 python tools/timeline.py trains 20 > "$OUT/timeline_unfused_late_wb_trigger.txt" 2>&1; echo "timeline exit $?"; sed -n 20,60p "$OUT/timeline_unfused_late_wb_trigger.txt"
timeout 300 python -m pytest tests/test_trains_gpu.py tests/test_step_gpu.py tests/test_rows_gpu.py -m gpu -q > "$OUT/pytest_some.log" 2>&1
echo "pytest some exit $?"; tail -3 "$OUT/pytest_some.log"
ls "$OUT"
