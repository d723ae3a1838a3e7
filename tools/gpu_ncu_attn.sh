#!/bin/bash
# ncu --set full (with source-level stall samples) of the tcgen05 triangle-attention kernel
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:pair_attention_tc5 -s 2 -c 1 -f -o gpurun_out/attn_tc5_${1:-x} \
  python tools/bench_attention.py > gpurun_out/ncu_attn_${1:-x}.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_attn_${1:-x}.log
