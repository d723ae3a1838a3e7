#!/bin/bash
# parity tests + short bench + ncu launch list; TAG names the outputs
TAG=${1:-x}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_$TAG.log
timeout 120 python tools/bench_ipa.py --B 4 --N 350 > gpurun_out/bench_ipa_$TAG.log 2>&1
cat gpurun_out/bench_ipa_$TAG.log
timeout 200 python tools/profile_step.py > gpurun_out/profile_$TAG.log 2>&1; head -14 gpurun_out/profile_$TAG.log; grep wall_ms gpurun_out/profile_$TAG.log
timeout 600 python bench.py --steps 1 --warmup 1 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json | cut -c1-1800; tail -5 gpurun_out/bench_$TAG.err
