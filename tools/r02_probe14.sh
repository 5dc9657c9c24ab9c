#!/bin/bash
# round-2 probe 14: three TMEM accumulators in the tcgen05 dense layer -- parity and timing
set -u
OUT=gpurun_out/r02p14
mkdir -p "$OUT"
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > "$OUT/pytest_gemm.log" 2>&1
echo "pytest gemm exit $?"; tail -3 "$OUT/pytest_gemm.log" | cut -c1-220
python - <<'PY' > "$OUT/gemm_time.txt" 2>&1
import sys, numpy as np, torch
sys.path.insert(0, ".")
from stochastic_gcn_b200 import ops, nn
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
n_nodes, k, n = 232965, 1204, 1501
feats = torch.randn((n_nodes, k), generator=gen, device=dev)
w = torch.randn((k, 128), generator=gen, device=dev) / np.sqrt(k)
packed = ops.pack_dense_weights(w)
idx = torch.randint(0, n_nodes, (n,), generator=gen, device=dev, dtype=torch.int32)
out = torch.empty((n, 128), device=dev); x0 = torch.empty((n, k), device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def fused(): ops.gathered_dense(feats, idx, packed, k, epilogue="ln_relu", out=out)
def unfused():
    ops.gather_rows(feats, idx, out=x0); nn.layer_norm_act(torch.mm(x0, w), None, None, 1e-9, True)
for name, fn in (("fused", fused), ("unfused", unfused), ("fused", fused), ("unfused", unfused)):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0.record()
    for _ in range(100): fn()
    e1.record(); torch.cuda.synchronize()
    print(name, "%.2f us" % (1e3 * e0.elapsed_time(e1) / 100))
PY
cat "$OUT/gemm_time.txt"
