#!/bin/bash
# round-2 final evidence: full GPU parity suite, bench lines, kernel micro-benchmarks, launch list, IPA layer-call ncu, same-seed report
TAG=${1:-r02f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -10 gpurun_out/pytest_$TAG.log
bash tools/gpu_r2_bench.sh $TAG
timeout 100 python tools/bench_gemm.py > gpurun_out/bench_gemm_$TAG.jsonl 2>&1; echo "bench_gemm rc=$?"
timeout 100 python tools/bench_attention.py > gpurun_out/bench_attention_$TAG.jsonl 2>&1; echo "bench_attention rc=$?"; cat gpurun_out/bench_attention_$TAG.jsonl
timeout 200 python tools/bench_ipa.py --B 8 --N 350 --iters 20 > gpurun_out/bench_ipa_$TAG.log 2>&1; echo "bench_ipa rc=$?"; tail -3 gpurun_out/bench_ipa_$TAG.log
bash tools/gpu_r2_profiles.sh $TAG
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pair_attention_tc5 -s 2 -c 1 -f -o gpurun_out/attn_tc5_$TAG \
  python tools/bench_attention.py > /dev/null 2>&1; echo "ncu attention rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:gemm_tf32x3 -f -o gpurun_out/gemm_$TAG \
  python tools/gemm_one.py 980000x768x192x128 980000x192x768x128 > /dev/null 2>&1; echo "ncu gemm rc=$?"
