#!/bin/bash
# round-2 probe 15: host-buffer leg through a device staging ring (D2D then D2H)
set -u
OUT=gpurun_out/r02p15
mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_trains_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > "$OUT/pytest.log" 2>&1
echo "pytest exit $?"; tail -3 "$OUT/pytest.log" | cut -c1-220
b() { name=$1; shift; timeout 300 python bench.py "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"; echo "$name exit $?"; python - "$OUT/$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ms/step %.5f  e2e %.5f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]))
except Exception as e: print("   parse failed", e)
PY
tail -3 "$OUT/$name.err"; }
b k20 --steps 20 --warmup 5 --no-cpu --no-also
b k2048 --steps 2048 --warmup 5 --no-cpu --no-also
b k20b --steps 20 --warmup 5 --no-cpu --no-also
