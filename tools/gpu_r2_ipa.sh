#!/bin/bash
# round 2: fused IPA kernel — parity, layer-call timings fused vs two-kernel
TAG=${1:-r02b}
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_ipa.py -x -q > gpurun_out/pytest_ipa_$TAG.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_ipa_$TAG.log
for B in 8 4 1; do
  timeout 120 python tools/bench_ipa.py --B $B --N 350 --profile 1 --graph 8 > gpurun_out/bench_ipa_fused_B${B}_$TAG.log 2>&1; tail -12 gpurun_out/bench_ipa_fused_B${B}_$TAG.log
done
timeout 120 python tools/bench_ipa.py --B 8 --N 262 --graph 8 > gpurun_out/bench_ipa_fused_B8_N262_$TAG.log 2>&1; tail -1 gpurun_out/bench_ipa_fused_B8_N262_$TAG.log
