#!/bin/bash
# parity tests + short bench + ncu launch list; TAG names the outputs
TAG=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_$TAG.log
for B in 4 8; do timeout 120 python tools/bench_ipa.py --B $B --N 350 >> gpurun_out/bench_ipa_$TAG.log 2>&1; done
cat gpurun_out/bench_ipa_$TAG.log
timeout 900 python bench.py --steps 1 --warmup 1 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json | cut -c1-1800; tail -5 gpurun_out/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 1 --num-t 2 --no-cpu-baseline > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu rc=$?"
