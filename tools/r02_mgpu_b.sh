#!/bin/bash
# multi-GPU visit: claims off the chain (split apply) -- parity of the trains forms on every GPU, A/B bench, timeline
set -u
N=$(nvidia-smi -L | wc -l)
OUT=gpurun_out/r02mgpu${N}c
mkdir -p "$OUT"
rm -f gpurun_out/mgpu_check_$N.log gpurun_out/mgpu_check_fullsize.log
timeout 900 python -m pytest tests/test_sharding_gpu.py -q -k "all_ranks or (benched and cv-replicated)" > "$OUT/pytest.log" 2>&1; echo "pytest exit $?"; tail -6 "$OUT/pytest.log"
cp gpurun_out/mgpu_check_$N.log gpurun_out/mgpu_check_fullsize.log "$OUT/" 2>/dev/null
b() { name=$1; port=$2; shift 2; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"; echo "$name exit $?"; python - "$OUT/$name.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("   ms/step %.5f  e2e %.5f  value %.4g" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"]), {k:(v.get("ms_per_step"), v.get("error")) for k,v in d.get("also",{}).items()})
except Exception as e: print("   parse failed", e)
PY
tail -2 "$OUT/$name.err" | cut -c1-300; }
b k20 29601 --steps 20 --warmup 5 --no-also
b k2000 29602 --steps 2000 --warmup 5 --no-also
SGCN_WB_SPLIT=0 b k20_nosplit 29603 --steps 20 --warmup 5 --no-also
SGCN_WB_SPLIT=0 b k2000_nosplit 29605 --steps 2000 --warmup 5 --no-also
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29604 tools/timeline_mgpu.py 20 > "$OUT/timeline.txt" 2> "$OUT/timeline.err"; echo "timeline exit $?"; sed -n 14,60p "$OUT/timeline.txt"; tail -3 "$OUT/timeline.err"
ls "$OUT"
