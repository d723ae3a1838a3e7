#!/bin/bash
# ncu --set full capture of one IPA layer-call (B=$1, N=350): every abx kernel of the 4th timed call
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ipa_|gemm_tf32x3' -s 25 -c 6 -f -o gpurun_out/ipa_r01b \
  python tools/bench_ipa.py --B ${1:-8} --N 350 --iters 3 > gpurun_out/ncu_ipa.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/ncu_ipa.log
