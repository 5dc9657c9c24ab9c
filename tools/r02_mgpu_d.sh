#!/bin/bash
# 2-GPU A/B: stream priorities and programmatic launches in the multi-GPU chain
set -u
N=$(nvidia-smi -L | wc -l)
OUT=gpurun_out/r02mgpu${N}e
mkdir -p "$OUT"
b() { name=$1; port=$2; shift 2; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"; echo "$name exit $?"; python - "$OUT/$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ms/step %.5f  e2e %.5f  value %.4g" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"]))
except Exception as e: print("   parse failed", e)
PY
tail -2 "$OUT/$name.err" | cut -c1-300; }
SGCN_STEP_PRIORITY=2 b k2000_prio2 29602 --steps 2000 --warmup 5 --no-also
SGCN_STEP_PRIORITY=1 b k2000_prio1 29603 --steps 2000 --warmup 5 --no-also
SGCN_PDL=0 b k2000_nopdl 29605 --steps 2000 --warmup 5 --no-also
SGCN_PDL=0 SGCN_STEP_PRIORITY=2 b k2000_nopdl_prio2 29606 --steps 2000 --warmup 5 --no-also
SGCN_STEP_PRIORITY=2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29604 tools/timeline_mgpu.py 20 > "$OUT/timeline_prio2.txt" 2> "$OUT/timeline.err"; echo "timeline exit $?"; sed -n 30,60p "$OUT/timeline_prio2.txt"
ls "$OUT"
