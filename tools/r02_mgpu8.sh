#!/bin/bash
# 4- / 8-GPU visit (charged N x): parity of the trains schedule on every GPU, bench at N = device count for the
# replicated and the sharded tables, per-rank timelines.  Kept short on purpose.
set -u
N=$(nvidia-smi -L | wc -l)
OUT=gpurun_out/r02mgpu$N
mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpus.txt"
chk() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 tests/mgpu_check.py peer $2 $3 $4 > "$OUT/check_$2_$3_$4.log" 2>&1; echo "check $2 $3 $4 exit $?"; tail -1 "$OUT/check_$2_$3_$4.log" | cut -c1-200; }
chk 29611 cv trains replicated
chk 29612 cvd trains-graph replicated
chk 29613 cv trains-graph sharded
b() { name=$1; port=$2; shift 2; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"; echo "$name exit $?"; cut -c1-260 "$OUT/$name.json"; tail -2 "$OUT/$name.err" | cut -c1-300; }
b k20 29621 --steps 20 --warmup 5
b k2000 29622 --steps 2000 --warmup 5 --no-also
b k200_sharded 29623 --steps 200 --warmup 5 --no-also --tables sharded
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29624 tools/timeline_mgpu.py 20 > "$OUT/timeline.txt" 2> "$OUT/timeline.err"; echo "timeline exit $?"; grep "^rank" "$OUT/timeline.txt"
ls "$OUT"
