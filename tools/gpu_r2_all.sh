#!/bin/bash
# round 2: full GPU parity suite + same-seed report + default bench
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_$TAG.log
timeout 900 python tools/same_seed.py --tag $TAG > gpurun_out/same_seed_$TAG.log 2>&1; echo "same_seed rc=$?"; tail -18 gpurun_out/same_seed_$TAG.log | cut -c1-250
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cut -c1-3000 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
