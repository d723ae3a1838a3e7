#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_gemm.py -x -q > gpurun_out/pytest_gemm.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gemm.log
ABX_GEMM_TRUST_TRUNC=0 timeout 100 python tools/gemm_trunc_probe.py /tmp/y0.pt; ABX_GEMM_TRUST_TRUNC=1 timeout 100 python tools/gemm_trunc_probe.py /tmp/y1.pt
python -c "import torch; a=torch.load('/tmp/y0.pt'); b=torch.load('/tmp/y1.pt'); print('trust_trunc bit-identical:', torch.equal(a,b), float((a-b).abs().max()))"
timeout 120 python tools/bench_gemm.py > gpurun_out/bench_gemm.log 2>&1; echo "bench rc=$?"
cat gpurun_out/bench_gemm.log | tail -12
