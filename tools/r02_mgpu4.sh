#!/bin/bash
# 4-GPU visit (charged 4 x, kept short): parity on every GPU, bench at N = 4 (final code and the previous push / trigger
# form), per-rank timelines
set -u
N=$(nvidia-smi -L | wc -l)
OUT=gpurun_out/r02mgpu$N
mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpus.txt"
chk() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 tests/mgpu_check.py peer $2 $3 $4 > "$OUT/check_$2_$3_$4.log" 2>&1; echo "check $2 $3 $4 exit $?"; grep "mgpu_check ok\|Error\|differs" "$OUT/check_$2_$3_$4.log" | head -3 | cut -c1-240; }
chk 29611 cvd trains-graph replicated
b() { name=$1; port=$2; shift 2; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"; echo "$name exit $?"; python - "$OUT/$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ms/step %.5f  e2e %.5f  value %.4g" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"]), {k:(v.get("ms_per_step"), v.get("value"), v.get("error")) for k,v in d.get("also",{}).items()})
except Exception as e: print("   parse failed", e)
PY
tail -2 "$OUT/$name.err" | cut -c1-300; }
b k20 29621 --steps 20 --warmup 5
b k2000 29622 --steps 2000 --warmup 5 --no-also
SGCN_WB_TRIGGER=0 SGCN_PUSH_BLOCKS=192 b k2000_old 29623 --steps 2000 --warmup 5 --no-also
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29624 tools/timeline_mgpu.py 20 > "$OUT/timeline.txt" 2> "$OUT/timeline.err"; echo "timeline exit $?"; grep "^rank" "$OUT/timeline.txt"
ls "$OUT"
