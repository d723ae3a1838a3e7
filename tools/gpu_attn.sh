#!/bin/bash
# triangle-attention check: parity tests of the three kernels + micro-benchmark (tc5 vs mma.sync)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_trunk_ops.py -q -x -k "pair_attention" 2>&1 | tail -8
timeout 120 python tools/bench_attention.py 2>&1 | tail -4 | tee gpurun_out/bench_attention_${1:-x}.log
