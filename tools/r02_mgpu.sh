#!/bin/bash
# multi-GPU visit: parity on every GPU of the box, bench at N = device count, per-rank timelines
set -u
N=$(nvidia-smi -L | wc -l)
OUT=gpurun_out/r02mgpu$N
mkdir -p "$OUT"
nvidia-smi -L > "$OUT/gpus.txt"
rm -f gpurun_out/mgpu_check_$N.log
timeout 900 python -m pytest tests/test_sharding_gpu.py -x -q > "$OUT/pytest.log" 2>&1; echo "pytest exit $?"; tail -6 "$OUT/pytest.log"
cp gpurun_out/mgpu_check_$N.log "$OUT/" 2>/dev/null
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
b() { name=$1; port=$2; shift 2; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"; echo "$name exit $?"; cut -c1-260 "$OUT/$name.json"; tail -3 "$OUT/$name.err"; }
b k20 29601 --steps 20 --warmup 5
b k2000 29602 --steps 2000 --warmup 5 --no-also
SGCN_WB_RING=0 b k2000_noring 29603 --steps 2000 --warmup 5 --no-also
b k2000_sharded 29605 --steps 2000 --warmup 5 --no-also --tables sharded
b k20_sharded 29606 --steps 20 --warmup 5 --no-also --tables sharded
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29604 tools/timeline_mgpu.py 20 > "$OUT/timeline.txt" 2> "$OUT/timeline.err"; echo "timeline exit $?"; head -50 "$OUT/timeline.txt"; tail -3 "$OUT/timeline.err"
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-also --no-cpu > "$OUT/n1_k20.json" 2> "$OUT/n1_k20.err"; cut -c1-260 "$OUT/n1_k20.json"
for m in late early noshare; do timeout 60 python tools/debug_train_wait.py $m > "$OUT/debug_train_wait_$m.txt" 2>&1; head -12 "$OUT/debug_train_wait_$m.txt"; done
ls -la "$OUT"
