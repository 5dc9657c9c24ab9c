#!/bin/bash
# One GPU-box visit: parity tests, smoke, both bench arms, ncu launch list, `ncu --set full` captures of the
# dominant kernel and of the tcgen05 dense layer.  Everything lands in gpurun_out/<tag>/ (copied to profiles/ by hand).
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
set -u
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/nvsmi.txt" 2>&1

timeout 200 python bench.py --steps 20 --warmup 5 > "$OUT/bench_n1_k20.json" 2> "$OUT/bench_n1_k20.err"; echo "bench k20 exit $?"
cut -c1-400 "$OUT/bench_n1_k20.json"
timeout 200 python bench.py --no-also > "$OUT/bench_n1.json" 2> "$OUT/bench_n1.err"; echo "bench default exit $?"
cut -c1-400 "$OUT/bench_n1.json"
if [ "${SKIP_REF:-0}" != 1 ]; then
  timeout 120 python bench.py --impl reference --steps 20 --warmup 5 > "$OUT/bench_ref.json" 2> "$OUT/bench_ref.err"
  echo "ref exit $?"; cut -c1-300 "$OUT/bench_ref.json"
fi

if [ "${SKIP_NCU:-0}" != 1 ]; then
  # launch list (cold-cache, serialised): the kernel's SHARE of the step must agree with the bench
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 300 -c 400 --csv \
      --log-file "$OUT/launches.csv" python bench.py --steps 40 --warmup 3 --no-cpu --no-also > "$OUT/ncu_launch_bench.log" 2>&1
  echo "ncu launches exit $?"
  # one full capture of the dominant kernel (3 launches, after the warm-up launches), programmatic launches ON
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:full_mean_kernel \
      --launch-skip 30 -c 3 -f -o "$OUT/full_mean" python bench.py --steps 20 --warmup 3 --no-cpu --no-also \
      > "$OUT/ncu_full_bench.log" 2>&1
  echo "ncu full exit $?"
  ncu -i "$OUT/full_mean.ncu-rep" --page raw --csv > "$OUT/full_mean_raw.csv" 2>/dev/null
  ncu -i "$OUT/full_mean.ncu-rep" --page source --csv > "$OUT/full_mean_source.csv" 2>/dev/null
  # the tcgen05 dense layer
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:gather_gemm \
      --launch-skip 3 -c 1 -f -o "$OUT/gather_gemm" python tools/gemm_probe.py > "$OUT/ncu_gemm.log" 2>&1
  echo "ncu gemm exit $?"
  ncu -i "$OUT/gather_gemm.ncu-rep" --page raw --csv > "$OUT/gather_gemm_raw.csv" 2>/dev/null
  rm -f "$OUT"/*.ncu-rep.tmp
fi

if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout 900 python -m pytest tests -m gpu -q > "$OUT/pytest_gpu.log" 2>&1
  echo "pytest exit $?" >> "$OUT/pytest_gpu.log"
  tail -3 "$OUT/pytest_gpu.log"
  timeout 200 python __graft_entry__.py --smoke > "$OUT/smoke.log" 2>&1; echo "smoke exit $?" >> "$OUT/smoke.log"
  tail -3 "$OUT/smoke.log"
  cp gpurun_out/parity_fullsize.json "$OUT/" 2>/dev/null
fi
ls -la "$OUT"
