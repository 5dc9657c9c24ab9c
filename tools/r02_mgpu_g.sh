#!/bin/bash
# 2-GPU bench line of the final code with the configs[3] / configs[4] extra keys
set -u
N=$(nvidia-smi -L | wc -l)
OUT=gpurun_out/r02mgpu${N}h
mkdir -p "$OUT"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 20 --warmup 5 > "$OUT/k20.json" 2> "$OUT/k20.err"; echo "bench exit $?"
python - "$OUT/k20.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step %.5f e2e %.5f value %.4g" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["value"]), {k:(v.get("ms_per_step"), v.get("error")) for k,v in d.get("also",{}).items()})
PY
