#!/bin/bash
# round-2 probe 17: the batch-32768 stress point of SURVEY 8d (general multi-kernel sampler path, native driver)
set -u
OUT=gpurun_out/r02p17
mkdir -p "$OUT"
timeout 200 python bench.py --workload reddit_cv_b32768 --steps 6 --warmup 3 --no-also --no-cpu > "$OUT/b32768.json" 2> "$OUT/b32768.err"; echo "exit $?"
tail -3 "$OUT/b32768.err" | cut -c1-300
python - "$OUT/b32768.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.5f e2e %.5f value %.4g driver %s kernel us %.1f frac %.3f" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"], d["schedule"]["driver"], d["roofline"]["us_per_launch"], d["roofline"]["frac"]))
PY
