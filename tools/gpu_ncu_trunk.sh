#!/bin/bash
# ncu --set full of the tcgen05 3xTF32 GEMM at the pair-stack shapes (M = 4 x 350^2 rows; 192 -> 768 and 768 -> 192);
# the triangle-attention capture of the same round used: -k regex:pair_attention_mma -s 3 -c 1 python tools/bench_attention.py
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 1 -c 1 -f -o gpurun_out/gemm_192_768 \
  python tools/gemm_one.py 490000x768x192x128 > gpurun_out/ncu_trunk.log 2>&1; echo "ncu gemm1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32x3 -s 1 -c 1 -f -o gpurun_out/gemm_768_192 \
  python tools/gemm_one.py 490000x192x768x128 >> gpurun_out/ncu_trunk.log 2>&1; echo "ncu gemm2 rc=$?"
for r in gemm_192_768 gemm_768_192; do
  timeout 120 ncu -i gpurun_out/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
  timeout 120 ncu -i gpurun_out/$r.ncu-rep --page details --csv > gpurun_out/${r}_details.csv 2>/dev/null
done
ls -la gpurun_out | tail -8
