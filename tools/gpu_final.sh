#!/bin/bash
# round-end validation: full GPU parity suite, default bench, ncu launch list of the bench command, ncu --set full of one IPA layer-call
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -14 gpurun_out/pytest_$TAG.log
timeout 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cut -c1-2400 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 1 --warmup 1 --num-t 2 --no-cpu-baseline > gpurun_out/bench_ncu_$TAG.log 2>&1; echo "ncu launches rc=$?"
python tools/summarize_launches.py gpurun_out/launches_$TAG.csv --top 14 | tee gpurun_out/launches_$TAG.md | head -24
bash tools/gpu_ncu_ipa.sh 8
timeout 120 ncu -i gpurun_out/ipa_r01b.ncu-rep --page raw --csv > gpurun_out/ipa_layer_call_raw_$TAG.csv 2>/dev/null; echo "raw csv rc=$?"
