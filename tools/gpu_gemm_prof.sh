#!/bin/bash
# per-role wait profile of the GEMM main loop: rebuild with the counters compiled in, run, rebuild the product library
mkdir -p gpurun_out
ABX_GEMM_PROFILE=1 python -m abx_b200.build > /dev/null 2>&1
for kb in 1 2; do echo "kb_per_drain=$kb"; ABX_GEMM_KB_PER_DRAIN=$kb ABX_GEMM_PROF=1 timeout 100 python tools/gemm_prof.py; done 2>&1 | tee gpurun_out/gemm_prof_${1:-x}.log
python -m abx_b200.build > /dev/null 2>&1
