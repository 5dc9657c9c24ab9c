#!/bin/bash
set -u
OUT=gpurun_out/${1:-mgpu}
mkdir -p "$OUT"
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_sharding_gpu.py -x -q > "$OUT/pytest.log" 2>&1; tail -5 "$OUT/pytest.log"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2000 --warmup 10 > "$OUT/bench_n2.json" 2> "$OUT/bench_n2.err"; echo "bench n2 exit $?"
cat "$OUT/bench_n2.json" | cut -c1-1500; tail -5 "$OUT/bench_n2.err"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 tools/timeline_mgpu.py 8 > "$OUT/timeline.txt" 2> "$OUT/timeline.err"; head -2 "$OUT/timeline.txt"; tail -24 "$OUT/timeline.txt"
