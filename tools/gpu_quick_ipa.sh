#!/bin/bash
# quick check of the fused IPA kernel: parity on a few sizes (watchdog record) + layer-call timing at B=8
TAG=${1:-q}
mkdir -p gpurun_out
timeout 120 python tools/ipa_debug.py 2,37 3,100 8,350 2>&1 | tail -14
timeout 120 python tools/bench_ipa.py --B 8 --N 350 --profile 1 --prof 1 --graph 8 > gpurun_out/bench_ipa_fused_B8_$TAG.log 2>&1; grep -v "^ *0.0 us\|warn" gpurun_out/bench_ipa_fused_B8_$TAG.log | tail -14
