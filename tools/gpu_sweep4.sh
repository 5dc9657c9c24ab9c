#!/bin/bash
set -u
OUT=gpurun_out/${1:-sweep5}
mkdir -p "$OUT"
export SGCN_PDL=1
timeout 300 python -m pytest tests/test_step_gpu.py tests/test_rows_gpu.py tests/test_aggregate_gpu.py -x -q > "$OUT/pytest.log" 2>&1; tail -2 "$OUT/pytest.log"
for M in 0 1; do
  if [ $M = 1 ]; then export SGCN_NO_MATCH_CARVEOUT=1; fi
  echo "== no-match-carveout $M"
  timeout 200 python bench.py --no-cpu --steps 4096 > "$OUT/bench_M$M.json" 2> "$OUT/bench_M$M.err"; echo "bench exit $?"
  python - "$OUT/bench_M$M.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("ms/step %.5f serial %.5f e2e %.5f kern_us %.2f frac %.3f" % (d["ms_per_step"], d["schedule"]["ms_per_step_one_graph_back_to_back"], d["e2e"]["ms_per_step"], d["roofline"]["us_per_launch"], d["roofline"]["frac"]))
except Exception as e:
    print("no bench line", e)
PY
  tail -2 "$OUT/bench_M$M.err"
done
unset SGCN_NO_MATCH_CARVEOUT
timeout 200 python tools/timeline.py pipelined 8 > "$OUT/timeline.txt" 2>&1; tail -26 "$OUT/timeline.txt"
