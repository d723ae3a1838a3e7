#!/bin/bash
# GEMM check: parity tests (GEMM + trunk ops that sit on it) and the micro-benchmark
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/pytest_gemm.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gemm.log
timeout 200 python -m pytest tests/test_gpu_trunk_ops.py -x -q -k "not pair_attention" 2>&1 | tail -5
timeout 100 python tools/bench_gemm.py > gpurun_out/bench_gemm_${1:-x}.log 2>&1; echo "bench rc=$?"
cat gpurun_out/bench_gemm_${1:-x}.log | tail -12
