#!/bin/bash
set -u
OUT=gpurun_out/${1:-pp}
mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_precompute_gpu.py -x -q > "$OUT/pytest.log" 2>&1; tail -4 "$OUT/pytest.log"
timeout 400 python tools/pp_bench.py > "$OUT/pp_bench.jsonl" 2> "$OUT/pp_bench.err"; echo "pp exit $?"; cat "$OUT/pp_bench.jsonl"; tail -3 "$OUT/pp_bench.err"
