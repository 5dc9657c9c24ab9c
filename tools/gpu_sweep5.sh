#!/bin/bash
set -u
OUT=gpurun_out/${1:-sweep6}
mkdir -p "$OUT"
export SGCN_PDL=1
timeout 400 python -m pytest tests -m gpu -x -q > "$OUT/pytest.log" 2>&1; tail -2 "$OUT/pytest.log"
for M in "" "CUDA_DEVICE_MAX_CONNECTIONS=32"; do
  echo "== env $M"
  env $M timeout 200 python bench.py --no-cpu --steps 4096 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $?"
  python - "$OUT/bench.json" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print("ms/step %.5f serial %.5f e2e %.5f kern_us %.2f frac %.3f launches/step %.1f" % (d["ms_per_step"], d["schedule"]["ms_per_step_one_graph_back_to_back"], d["e2e"]["ms_per_step"], d["roofline"]["us_per_launch"], d["roofline"]["frac"], d["launches_per_step"]))
except Exception as e:
    print("no bench line", e)
PY
  tail -2 "$OUT/bench.err"
done
timeout 200 python tools/timeline.py pipelined 8 > "$OUT/timeline.txt" 2>&1; tail -22 "$OUT/timeline.txt"
