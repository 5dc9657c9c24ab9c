#!/bin/bash
# 2-GPU A/B: copy pass of the exchange launched plainly vs programmatically
set -u
N=$(nvidia-smi -L | wc -l)
OUT=gpurun_out/r02mgpu${N}f
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
SGCN_WB_COPY_PDL=1 b k2000_copy_pdl 29603 --steps 2000 --warmup 5 --no-also
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29604 tools/timeline_mgpu.py 20 > "$OUT/timeline.txt" 2> "$OUT/timeline.err"; echo "timeline exit $?"; sed -n 30,56p "$OUT/timeline.txt"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/mgpu_check.py peer cv trains-graph replicated > "$OUT/check.log" 2>&1; echo "check exit $?"; grep "mgpu_check ok" "$OUT/check.log"
ls "$OUT"
