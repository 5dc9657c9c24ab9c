#!/bin/bash
# round-2 probe 12: final code -- the whole GPU suite, smoke, the driver's bench command line
set -u
OUT=gpurun_out/r02p12
mkdir -p "$OUT"
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -q -x > "$OUT/pytest_gpu.log" 2>&1
echo "pytest exit $? after $((SECONDS-T0)) s"; tail -6 "$OUT/pytest_gpu.log" | cut -c1-250
timeout 200 python __graft_entry__.py --smoke > "$OUT/smoke.log" 2>&1; echo "smoke exit $?"; tail -2 "$OUT/smoke.log"
T0=$SECONDS
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > "$OUT/bench_k20_full.json" 2> "$OUT/bench_k20_full.err"; echo "bench k20 full exit $? after $((SECONDS-T0)) s"
python - "$OUT/bench_k20_full.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.5f e2e %.5f launches %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches"]))
for k,v in d.get("also",{}).items(): print(k, {a:b for a,b in v.items() if a not in ("what","config","last_step_sizes")})
PY
cp gpurun_out/parity_fullsize.json "$OUT/" 2>/dev/null
ls "$OUT"
