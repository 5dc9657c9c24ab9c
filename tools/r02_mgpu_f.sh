#!/bin/bash
# 2-GPU check of the final code: multi-process oracle (eager trains incl. the cold first run, graphs, sharded), one bench line
set -u
N=$(nvidia-smi -L | wc -l)
OUT=gpurun_out/r02mgpu${N}g
mkdir -p "$OUT"
chk() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 tests/mgpu_check.py peer $2 $3 $4 > "$OUT/check_$2_$3_$4.log" 2>&1; echo "check $2 $3 $4 exit $?"; grep "mgpu_check ok\|Error\|differs" "$OUT/check_$2_$3_$4.log" | head -3 | cut -c1-240; }
chk 29611 cv trains replicated
chk 29612 cvd trains-graph sharded
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 20 --warmup 5 --no-also > "$OUT/k20.json" 2> "$OUT/k20.err"; echo "bench exit $?"
python - "$OUT/k20.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.5f e2e %.5f value %.4g" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"]))
PY
