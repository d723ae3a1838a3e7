#!/bin/bash
# ncu --set full capture of the IPA layer-call kernels (B=4, N=350), one launch of each after warm-up
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ipa_|linear_f32' -s 24 -c 8 -f -o gpurun_out/ipa_r01 \
  python tools/bench_ipa.py --B 4 --N 350 --iters 2 > gpurun_out/ncu_ipa.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_ipa.log
