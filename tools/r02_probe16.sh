#!/bin/bash
# round-2 probe 16: cold-process first run of the trains driver (no kernels loaded), then the whole GPU suite
set -u
OUT=gpurun_out/r02p16
mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_trains_gpu.py -m gpu -q -x > "$OUT/pytest_trains_cold.log" 2>&1
echo "pytest trains (cold process) exit $?"; tail -3 "$OUT/pytest_trains_cold.log" | cut -c1-220
timeout 300 python -m pytest tests/test_fullsize_gpu.py -m gpu -q -x -k "device_graph_of_20" > "$OUT/pytest_fullsize_cold.log" 2>&1
echo "pytest fullsize (cold process) exit $?"; tail -3 "$OUT/pytest_fullsize_cold.log" | cut -c1-220
T0=$SECONDS
timeout 900 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1
echo "pytest all exit $? after $((SECONDS-T0)) s"; tail -3 "$OUT/pytest_gpu.log" | cut -c1-220
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-also --no-cpu > "$OUT/bench_k20.json" 2> "$OUT/bench_k20.err"; echo "bench exit $?"
python - "$OUT/bench_k20.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.5f e2e %.5f" % (d["ms_per_step"], d["e2e"]["ms_per_step"]))
PY
