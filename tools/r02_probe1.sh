#!/bin/bash
# round-2 probe 1: what the existing drivers do on the driver's own command line, the unrun checks, the unrun workloads
set -u
OUT=gpurun_out/r02p1
mkdir -p "$OUT"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$OUT/nvsmi.txt" 2>&1
b() { name=$1; shift; timeout 300 python bench.py "$@" > "$OUT/$name.json" 2> "$OUT/$name.err"; echo "$name exit $?"; cut -c1-400 "$OUT/$name.json"; }
b graph_k20 --steps 20 --warmup 5 --no-cpu
b ahead_k20 --steps 20 --warmup 5 --no-cpu --driver ahead
b native_k20 --steps 20 --warmup 5 --no-cpu --driver native
b ahead_k2048 --steps 2048 --warmup 5 --no-cpu --driver ahead --steps-per-graph 32
b graph_k2048 --steps 2048 --warmup 5 --no-cpu
timeout 200 python tests/ahead_host_check.py > "$OUT/ahead_host_check.log" 2>&1; echo "ahead_host_check exit $?"; tail -3 "$OUT/ahead_host_check.log"
timeout 120 python tools/timeline.py ahead 16 > "$OUT/timeline_ahead.txt" 2>&1; echo "timeline exit $?"; head -60 "$OUT/timeline_ahead.txt"
b pubmed_cvd --workload pubmed_cvd --steps 200 --warmup 5
b reddit_cvd --workload reddit_cvd --steps 200 --warmup 5 --no-cpu
b powerlaw_ns --workload powerlaw_ns --steps 200 --warmup 5 --no-cpu
b reddit_b4096 --workload reddit_cv_b4096 --steps 64 --warmup 5 --no-cpu
b reddit_b32768 --workload reddit_cv_b32768 --steps 32 --warmup 5 --no-cpu
ls -la "$OUT"
