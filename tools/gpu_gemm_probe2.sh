#!/bin/bash
mkdir -p gpurun_out
for kb in 1 2 6; do for sk in 30 14 0; do echo -n "kb=$kb "; ABX_GEMM_KB_PER_DRAIN=$kb ABX_GEMM_DEBUG_SKIP=$sk timeout 60 python tools/gemm_probe.py; done; done 2>&1 | tee gpurun_out/gemm_probe2_${1:-x}.log
