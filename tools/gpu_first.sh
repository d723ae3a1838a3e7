#!/bin/bash
# first GPU pass: parity tests, micro-bench, bench line, ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for B in 1 4 8; do timeout 120 python tools/bench_ipa.py --B $B --N 350 >> gpurun_out/bench_ipa.log 2>&1; done
timeout 120 python tools/bench_ipa.py --B 4 --N 350 --precomputed-bias 0 >> gpurun_out/bench_ipa.log 2>&1
cat gpurun_out/bench_ipa.log
timeout 900 python bench.py --steps 1 --warmup 1 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err; echo "bench rc=$?"
cat gpurun_out/bench_first.json; tail -5 gpurun_out/bench_first.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 1 --num-t 2 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
