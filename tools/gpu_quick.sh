#!/bin/bash
TAG=${1:-q}
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/ipa_launches_$TAG.csv python tools/bench_ipa.py --B 4 --N 350 --iters 3 > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/ipa_launches_$TAG.csv --top 12
timeout 200 python tools/profile_step.py --linear-breakdown 0 > gpurun_out/profile_$TAG.log 2>&1; grep wall_ms gpurun_out/profile_$TAG.log; grep -A28 "Self CUDA %" gpurun_out/profile_$TAG.log | tail -27 | cut -c1-92,190-290
