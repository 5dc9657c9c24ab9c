#!/bin/bash
# last check of HEAD on a fresh box: smoke + the driver's two bench command lines
set -u
OUT=gpurun_out/r02p18
mkdir -p "$OUT"
T0=$SECONDS
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > "$OUT/smoke.log" 2>&1; echo "smoke exit $? after $((SECONDS-T0)) s"; tail -2 "$OUT/smoke.log"
T0=$SECONDS
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > "$OUT/bench.json" 2> "$OUT/bench.err"; echo "bench exit $? after $((SECONDS-T0)) s"
python - "$OUT/bench.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.5f e2e %.5f roofline %.3f/%.3f cpu %.3f ms" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["frac"], d["roofline"]["frac_dram"], d["cpu_baseline"]["ms_per_step"]))
print(sorted(d.keys()))
PY
