#!/bin/bash
# One GPU-box visit: parity tests, smoke, both bench arms, ncu launch list, one ncu --set full capture
# of the dominant kernel.  Everything lands in gpurun_out/ (copied to profiles/ by hand afterwards).
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
set -u
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/nvsmi.txt" 2>&1

if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout 300 python -m pytest tests -m gpu -x -q > "$OUT/pytest_gpu.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
  tail -3 "$OUT/pytest_gpu.log"
  timeout 200 python __graft_entry__.py --smoke > "$OUT/smoke.log" 2>&1; echo "smoke exit $?" >> "$OUT/smoke.log"
  tail -2 "$OUT/smoke.log"
fi

timeout 200 python bench.py > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; echo "bench exit $?"
cat "$OUT/bench_n1.json"
if [ "${SKIP_REF:-0}" != 1 ]; then
  timeout 120 python bench.py --impl reference --steps 40 --warmup 3 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"
  echo "ref exit $?"; cat "$OUT/bench_ref.json"
fi

if [ "${SKIP_NCU:-0}" != 1 ]; then
  # launch list (cold-cache, serialised): the kernel's SHARE of the step must agree with the bench
  SGCN_PDL=0 timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 200 -c 400 --csv \
      --log-file "$OUT/launches.csv" python bench.py --steps 32 --warmup 3 --no-cpu > "$OUT/ncu_launch_bench.log" 2>&1
  echo "ncu launches exit $?"
  # one full capture of the dominant kernel (3 launches, after the warm-up launches)
  SGCN_PDL=0 timeout 150 ncu --set full --clock-control none --import-source on -k regex:full_mean_kernel \
      --launch-skip 12 -c 3 -f -o "$OUT/full_mean" python bench.py --steps 16 --warmup 3 --no-cpu \
      > "$OUT/ncu_full_bench.log" 2>&1
  echo "ncu full exit $?"
  ncu -i "$OUT/full_mean.ncu-rep" --page raw --csv > "$OUT/full_mean_raw.csv" 2>/dev/null
fi
ls -la "$OUT"
