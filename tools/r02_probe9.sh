#!/bin/bash
# round-2 probe 9: 3 thread blocks per SM (80 registers); the whole GPU suite; the other workloads' bench lines
set -u
OUT=gpurun_out/r02p9
mkdir -p "$OUT"
b() { name=$1; shift; timeout 400 python bench.py "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"; echo "$name exit $?"; python - "$OUT/$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ms/step %.5f  e2e %.5f  kernel us %.2f  frac %.3f  value %.4g" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["us_per_launch"], d["roofline"]["frac"], d["value"]))
except Exception as e: print("   parse failed", e)
PY
tail -3 "$OUT/$name.err"; }
for K in 20 2048; do
b k${K} --steps $K --warmup 5 --no-cpu --no-also
SGCN_FULL_REGS=80 b k${K}_r80 --steps $K --warmup 5 --no-cpu --no-also
done
b pubmed_cvd --workload pubmed_cvd --steps 200 --warmup 5 --no-also
b reddit_cvd --workload reddit_cvd --steps 200 --warmup 5 --no-also --no-cpu
b powerlaw_ns --workload powerlaw_ns --steps 200 --warmup 5 --no-also --no-cpu
b reddit_cv_b4096 --workload reddit_cv_b4096 --steps 40 --warmup 3 --no-also --no-cpu
b reddit_cv_b32768 --workload reddit_cv_b32768 --steps 6 --warmup 3 --no-also --no-cpu
timeout 900 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1
echo "pytest gpu exit $?"; tail -8 "$OUT/pytest_gpu.log"
cp gpurun_out/parity_fullsize.json "$OUT/" 2>/dev/null
ls "$OUT"
