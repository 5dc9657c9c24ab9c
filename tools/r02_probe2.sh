#!/bin/bash
# round-2 probe 2: the trains schedule -- parity first, then timing on the driver's own command line
set -u
OUT=gpurun_out/r02p2
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/nvsmi.txt" 2>&1
timeout 300 python -m pytest tests/test_sampler_train_gpu.py tests/test_trains_gpu.py -x -q > "$OUT/pytest_new.log" 2>&1
echo "pytest new exit $?"; tail -15 "$OUT/pytest_new.log"
timeout 300 python __graft_entry__.py --smoke > "$OUT/smoke.log" 2>&1; echo "smoke exit $?"; tail -3 "$OUT/smoke.log"
b() { name=$1; shift; timeout 300 python bench.py "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"; echo "$name exit $?"; cut -c1-300 "$OUT/$name.json"; tail -3 "$OUT/$name.err"; }
b k20 --steps 20 --warmup 5 --no-cpu --no-also
b k20_overlap --steps 20 --warmup 5 --no-cpu --no-also --overlap-write-back
b k20_persistent --steps 20 --warmup 5 --no-cpu --no-also --persistent
b k20_persistent_ft2 --steps 20 --warmup 5 --no-cpu --no-also --persistent --first-train 2
b k2048_persistent --steps 2048 --warmup 5 --no-cpu --no-also --persistent
b k2048 --steps 2048 --warmup 5 --no-cpu --no-also
b k2048_overlap --steps 2048 --warmup 5 --no-cpu --no-also --overlap-write-back
b k2048_eager_persistent --steps 2048 --warmup 5 --no-cpu --no-also --eager-trains --persistent
SGCN_PDL=0 b k2048_nopdl --steps 2048 --warmup 5 --no-cpu --no-also
for m in "0 0" "1 0" "1 1"; do set -- $m
  OVERLAP=$1 PERSISTENT=$2 FIRST_TRAIN=4 timeout 120 python tools/timeline.py trains 20 > "$OUT/timeline_trains_ov$1_p$2.txt" 2>&1; echo "timeline exit $?"; head -3 "$OUT/timeline_trains_ov$1_p$2.txt"
done
timeout 600 python -m pytest tests/test_fullsize_gpu.py -x -q > "$OUT/pytest_fullsize.log" 2>&1
echo "pytest fullsize exit $?"; tail -15 "$OUT/pytest_fullsize.log"
b full_default --steps 20 --warmup 5
b pubmed_cvd --workload pubmed_cvd --steps 200 --warmup 5 --no-also --no-cpu
b pubmed_cvd_persistent --workload pubmed_cvd --steps 200 --warmup 5 --no-also --no-cpu --persistent
b reddit_cvd_persistent --workload reddit_cvd --steps 200 --warmup 5 --no-also --no-cpu --persistent
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_fullsize_gpu.py > "$OUT/pytest_gpu.log" 2>&1
echo "pytest gpu exit $?"; tail -8 "$OUT/pytest_gpu.log"
ls -la "$OUT"
