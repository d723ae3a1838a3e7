#!/bin/bash
# GEMM iteration: parity tests, micro-benchmark, per-role profile
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -4
timeout 200 python -m pytest tests/test_gpu_trunk_ops.py -x -q -k "not pair_attention" 2>&1 | tail -3
timeout 100 python tools/bench_gemm.py > gpurun_out/bench_gemm_${1:-x}.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
for l in open('gpurun_out/bench_gemm_${1:-x}.log'):
    try: r = json.loads(l)
    except Exception: print(l.strip()); continue
    print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in r.items() if k in ('M','N','K','abx_bn128_ms','abx_bn64_ms','abx_tflops_fp32_equiv','abx_gate_res_ms','abx_res_ms','abx_maxerr','torch_maxerr')})
PY
ABX_GEMM_PROFILE=1 python -m abx_b200.build > /dev/null 2>&1
ABX_GEMM_PROF=1 timeout 100 python tools/gemm_prof.py 2>&1 | tee gpurun_out/gemm_prof_${1:-x}.log
python -m abx_b200.build > /dev/null 2>&1
