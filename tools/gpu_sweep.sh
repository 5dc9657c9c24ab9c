#!/bin/bash
# GPU visit: det-dropout parity tests, memcheck of the bulk-copy full-mean on a small graph, A/B sweep.
set -u
OUT=gpurun_out/${1:-sweep}
mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_aggregate_gpu.py -x -q -k "det or full_mean" > "$OUT/pytest_det.log" 2>&1; tail -5 "$OUT/pytest_det.log"
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/full_mean_sweep.py --scale 0.03 --n 3 \
    --configs "1,12,16,2;1,8,32,1;1,16,5,3" > "$OUT/memcheck.log" 2>&1; echo "memcheck exit $?"; tail -8 "$OUT/memcheck.log"
timeout 500 python tools/full_mean_sweep.py ${SWEEP_ARGS:-} > "$OUT/sweep.jsonl" 2> "$OUT/sweep.err"; echo "sweep exit $?"
cat "$OUT/sweep.jsonl"; tail -3 "$OUT/sweep.err"
