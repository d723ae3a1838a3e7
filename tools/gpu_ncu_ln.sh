#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/ln_one.py 2>&1 | tail -2
timeout 200 ncu --set full --clock-control none --import-source on -k regex:layernorm_kernel -s 4 -c 1 -f -o gpurun_out/layernorm python tools/ln_one.py > gpurun_out/ncu_ln.log 2>&1; echo "ncu rc=$?"
timeout 100 ncu -i gpurun_out/layernorm.ncu-rep --page raw --csv > gpurun_out/layernorm_raw.csv 2>/dev/null
timeout 100 ncu -i gpurun_out/layernorm.ncu-rep --page details --csv > gpurun_out/layernorm_details.csv 2>/dev/null
