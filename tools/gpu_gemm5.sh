#!/bin/bash
# GEMM iteration: parity tests, micro-benchmark, per-shape timing table of a short sampler run
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -3
timeout 200 python -m pytest tests/test_gpu_trunk_ops.py -x -q -k "not pair_attention" 2>&1 | tail -3
timeout 100 python tools/bench_gemm.py > gpurun_out/bench_gemm_${1:-x}.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
for l in open('gpurun_out/bench_gemm_${1:-x}.log'):
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k in ('M','N','K','abx_bn128_ms','abx_bn64_ms','abx_tflops_fp32_equiv','abx_gate_res_ms','abx_res_ms')})
PY
ABX_GEMM_TIMING=1 timeout 400 python bench.py --steps 1 --warmup 1 --num-t 3 --cuda-graph 0 --no-cpu-baseline 2>gpurun_out/gemm_timing_${1:-x}.txt | cut -c1-120
grep -A45 "abx_gemm_tf32x3 timing" gpurun_out/gemm_timing_${1:-x}.txt | grep -E "timing|M=980000|M=350 N=350"
