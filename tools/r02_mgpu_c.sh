#!/bin/bash
# multi-GPU visit: late programmatic trigger of the copy pass + small push grid, A/B against the previous form
set -u
N=$(nvidia-smi -L | wc -l)
OUT=gpurun_out/r02mgpu${N}d
mkdir -p "$OUT"
b() { name=$1; port=$2; shift 2; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"; echo "$name exit $?"; python - "$OUT/$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ms/step %.5f  e2e %.5f  value %.4g" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"]))
except Exception as e: print("   parse failed", e)
PY
tail -2 "$OUT/$name.err" | cut -c1-300; }
b k2000 29602 --steps 2000 --warmup 5 --no-also
SGCN_WB_TRIGGER=0 SGCN_PUSH_BLOCKS=192 b k2000_old 29603 --steps 2000 --warmup 5 --no-also
SGCN_WB_TRIGGER=0 b k2000_early_push64 29605 --steps 2000 --warmup 5 --no-also
SGCN_PUSH_BLOCKS=192 b k2000_late_push192 29606 --steps 2000 --warmup 5 --no-also
SGCN_PUSH_BLOCKS=24 b k2000_late_push24 29607 --steps 2000 --warmup 5 --no-also
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29604 tools/timeline_mgpu.py 20 > "$OUT/timeline.txt" 2> "$OUT/timeline.err"; echo "timeline exit $?"; sed -n 14,50p "$OUT/timeline.txt"; tail -3 "$OUT/timeline.err"
timeout 300 python -m pytest tests/test_sharding_gpu.py -q -k "all_ranks" > "$OUT/pytest.log" 2>&1; echo "pytest exit $?"; tail -3 "$OUT/pytest.log"
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q > "$OUT/pytest_gemm.log" 2>&1; echo "pytest gemm exit $?"; tail -12 "$OUT/pytest_gemm.log" | cut -c1-200
ls "$OUT"
