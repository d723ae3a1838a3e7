#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -f -o gpurun_out/gemm_r01 \
  python tools/gemm_one.py 490000x768x192x128 4096x4096x4096x128 1400x256x2112x32 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_gemm.log
