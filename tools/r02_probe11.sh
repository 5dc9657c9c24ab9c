#!/bin/bash
# round-2 probe 11: the driver's own bench command line, timed (must finish within minutes)
set -u
OUT=gpurun_out/r02p11
mkdir -p "$OUT"
T0=$SECONDS
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > "$OUT/bench_k20_full.json" 2> "$OUT/bench_k20_full.err"; echo "bench k20 full exit $? after $((SECONDS-T0)) s"
tail -5 "$OUT/bench_k20_full.err"
python - "$OUT/bench_k20_full.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.5f e2e %.5f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]))
for k,v in d.get("also",{}).items(): print(k, {a:b for a,b in v.items() if a not in ("what","config","last_step_sizes")})
print(d.get("cpu_baseline",{}).get("ms_per_step"))
PY
T0=$SECONDS
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"; echo "ref exit $? after $((SECONDS-T0)) s"
ls "$OUT"
