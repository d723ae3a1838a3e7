#!/bin/bash
# timing probes of the GEMM main loop: ABX_GEMM_DEBUG_SKIP bits 2 = no conversion, 4 = no MMA, 8 = no drain, 16 = no TMA
mkdir -p gpurun_out
for sk in 0 16 2 18 4 8 30; do ABX_GEMM_DEBUG_SKIP=$sk timeout 60 python tools/gemm_probe.py; done 2>&1 | tee gpurun_out/gemm_probe_${1:-x}.log
