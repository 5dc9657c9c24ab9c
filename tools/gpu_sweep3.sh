#!/bin/bash
set -u
OUT=gpurun_out/${1:-sweep4}
mkdir -p "$OUT"
SGCN_PDL=1 timeout 300 python -m pytest tests/test_step_gpu.py tests/test_sampler_gpu.py tests/test_configs_gpu.py -x -q > "$OUT/pytest.log" 2>&1; tail -3 "$OUT/pytest.log"
for S in 8 16 32 64; do
  echo "== steps/graph $S"
  SGCN_PDL=1 timeout 200 python bench.py --no-cpu --steps 4096 --steps-per-graph $S > "$OUT/bench_S$S.json" 2> "$OUT/bench_S$S.err"; echo "bench exit $?"
  python - "$OUT/bench_S$S.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("ms/step %.5f serial %.5f e2e %.5f kern_us %.2f frac %.3f" % (d["ms_per_step"], d["schedule"]["ms_per_step_one_graph_back_to_back"], d["e2e"]["ms_per_step"], d["roofline"]["us_per_launch"], d["roofline"]["frac"]))
except Exception as e:
    print("no bench line", e)
PY
  tail -2 "$OUT/bench_S$S.err"
done
SGCN_PDL=1 timeout 200 python tools/timeline.py pipelined 16 > "$OUT/timeline.txt" 2>&1; tail -40 "$OUT/timeline.txt"
