#!/bin/bash
# ncu --set full of the fused IPA kernel (B=$1, N=350), one launch after warm-up
TAG=${2:-r02}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ipa_fused' -s 4 -c 1 -f -o gpurun_out/fused_$TAG \
  python tools/bench_ipa.py --B ${1:-8} --N 350 --iters 3 > gpurun_out/ncu_fused_$TAG.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/ncu_fused_$TAG.log
