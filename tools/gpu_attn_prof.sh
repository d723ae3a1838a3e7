#!/bin/bash
mkdir -p gpurun_out
ABX_ATTN_PROFILE=1 python -m abx_b200.build > /dev/null 2>&1
ABX_ATTN_PROF=1 timeout 120 python tools/bench_attention.py 2>&1 | tail -4 | tee gpurun_out/attn_prof_${1:-x}.log
python -m abx_b200.build > /dev/null 2>&1
