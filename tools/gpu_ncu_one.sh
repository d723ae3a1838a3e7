#!/bin/bash
# usage: gpu_ncu_one.sh <kernel regex> <skip> <tag>
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c 1 -f -o gpurun_out/one_$3 \
  python tools/bench_ipa.py --B 4 --N 350 --iters 3 > gpurun_out/ncu_one.log 2>&1; echo "ncu rc=$?"
