#!/bin/bash
# round 2: fused IPA kernel — parity, L2 probe, layer-call timings fused vs two-kernel
TAG=${1:-r02a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ipa.py -x -q > gpurun_out/pytest_ipa_$TAG.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_ipa_$TAG.log
timeout 120 tools/l2_probe.bin > gpurun_out/l2_probe_$TAG.jsonl 2>&1; cat gpurun_out/l2_probe_$TAG.jsonl
for B in 8 4 1; do
  timeout 200 python tools/bench_ipa.py --B $B --N 350 --profile 1 --graph 8 > gpurun_out/bench_ipa_fused_B${B}_$TAG.log 2>&1; cat gpurun_out/bench_ipa_fused_B${B}_$TAG.log
done
ABX_IPA_FUSED=0 timeout 200 python tools/bench_ipa.py --B 8 --N 350 --profile 1 --graph 8 > gpurun_out/bench_ipa_unfused_B8_$TAG.log 2>&1; cat gpurun_out/bench_ipa_unfused_B8_$TAG.log
timeout 200 python tools/bench_ipa.py --B 8 --N 262 --graph 8 > gpurun_out/bench_ipa_fused_B8_N262_$TAG.log 2>&1; cat gpurun_out/bench_ipa_fused_B8_N262_$TAG.log
