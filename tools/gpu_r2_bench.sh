#!/bin/bash
# round 2 bench lines: default (N=350 design), N=262 real-size stand-in, optimize / cdrs workload (BASELINE config 4)
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_$TAG.json
timeout 900 python bench.py --n-antigen 32 --no-cpu-baseline > gpurun_out/bench_n262_$TAG.json 2> gpurun_out/bench_n262_$TAG.err; echo "bench n262 rc=$?"; cut -c1-400 gpurun_out/bench_n262_$TAG.json
timeout 900 python bench.py --mode optimize --generate-area cdrs --optimize-steps 20 --no-cpu-baseline > gpurun_out/bench_optimize_$TAG.json 2> gpurun_out/bench_optimize_$TAG.err; echo "bench optimize rc=$?"; cut -c1-400 gpurun_out/bench_optimize_$TAG.json; tail -3 gpurun_out/bench_optimize_$TAG.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_$TAG.json 2>/dev/null; echo "reference rc=$?"; cut -c1-300 gpurun_out/bench_reference_$TAG.json
