#!/bin/bash
# round-2 probe 13: ncu --set full of the final (warp-specialised) tcgen05 dense layer
set -u
OUT=gpurun_out/r02p13
mkdir -p "$OUT"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gather_gemm \
    --launch-skip 3 -c 1 -f -o "$OUT/gather_gemm" python tools/gemm_probe.py > "$OUT/ncu_gemm.log" 2>&1
echo "ncu gemm exit $?"; tail -3 "$OUT/ncu_gemm.log"
ncu -i "$OUT/gather_gemm.ncu-rep" --page raw --csv > "$OUT/gather_gemm_raw.csv" 2>/dev/null
ncu -i "$OUT/gather_gemm.ncu-rep" --page source --csv > "$OUT/gather_gemm_source.csv" 2>/dev/null
ls -la "$OUT"
