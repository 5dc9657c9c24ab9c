#!/bin/bash
set -u
OUT=gpurun_out/${1:-sweep2}
mkdir -p "$OUT"
timeout 300 python tools/full_mean_sweep.py --configs "0;2" > "$OUT/sweep.jsonl" 2> "$OUT/sweep.err"; echo "sweep exit $?"; cat "$OUT/sweep.jsonl"; tail -3 "$OUT/sweep.err"
for cfg in "0 1" "2 0" "2 1"; do
  set -- $cfg
  echo "== variant $1 pdl $2"
  SGCN_FULL_VARIANT=$1 SGCN_PDL=$2 timeout 200 python -m pytest tests/test_step_gpu.py tests/test_aggregate_gpu.py -x -q > "$OUT/pytest_v$1_p$2.log" 2>&1; tail -2 "$OUT/pytest_v$1_p$2.log"
  SGCN_FULL_VARIANT=$1 SGCN_PDL=$2 timeout 200 python bench.py --no-cpu --steps 2000 > "$OUT/bench_v$1_p$2.json" 2> "$OUT/bench_v$1_p$2.err"; echo "bench exit $?"
  python - "$OUT/bench_v$1_p$2.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("ms/step %.5f serial %.5f e2e %.5f kern_us %.2f frac %.3f" % (d["ms_per_step"], d["schedule"]["ms_per_step_one_graph_back_to_back"], d["e2e"]["ms_per_step"], d["roofline"]["us_per_launch"], d["roofline"]["frac"]))
except Exception as e:
    print("no bench line", e)
PY
  tail -2 "$OUT/bench_v$1_p$2.err"
done
