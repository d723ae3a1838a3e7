#!/bin/bash
# round 2 evidence: ncu --set full of one IPA layer-call (+ the pair-bias pass), launch list of the bench command, same-seed report
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ipa_|gemm_tf32x3' -s 30 -c 5 -f -o gpurun_out/ipa_layer_call_$TAG \
  python tools/bench_ipa.py --B 8 --N 350 --iters 3 > gpurun_out/ncu_ipa_$TAG.log 2>&1; echo "ncu layer-call rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:'pair_bias' -s 3 -c 1 -f -o gpurun_out/ipa_pair_bias_$TAG \
  python tools/bench_ipa.py --B 8 --N 350 --iters 2 > gpurun_out/ncu_bias_$TAG.log 2>&1; echo "ncu bias rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 1 --warmup 1 --num-t 2 --no-cpu-baseline > gpurun_out/bench_ncu_$TAG.log 2>&1; echo "ncu launches rc=$?"
python tools/summarize_launches.py gpurun_out/launches_$TAG.csv --top 16 | tee gpurun_out/launches_$TAG.md | head -24
timeout 900 python tools/same_seed.py --tag $TAG > gpurun_out/same_seed_$TAG.log 2>&1; echo "same_seed rc=$?"; tail -16 gpurun_out/same_seed_$TAG.log | cut -c1-250
