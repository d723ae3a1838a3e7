#!/bin/bash
# A/B of the attention || pair-aggregation software pipeline (ABX_IPA_OVERLAP) inside the IPA layer-call (B200, N=350)
mkdir -p gpurun_out
OUT=gpurun_out/ipa_overlap_ab.jsonl; : > $OUT
timeout 300 python -m pytest tests/test_gpu_ipa.py -x -q 2>&1 | tail -3
for cfg in "8 0" "8 2" "8 3" "8 4" "4 0" "4 2" "4 4" "16 0" "16 4"; do
  set -- $cfg
  ABX_IPA_OVERLAP=$2 timeout 120 python tools/bench_ipa.py --B $1 --N 350 --iters 30 --graph 8 >> $OUT 2>gpurun_out/ipa_overlap_ab.err || tail -3 gpurun_out/ipa_overlap_ab.err
done
python - <<'PY'
import json
for l in open('gpurun_out/ipa_overlap_ab.jsonl'):
    d = json.loads(l)
    print(d['B'], 'overlap', d['overlap'], 'eager us %.1f frac %.3f | graph us %.1f frac %.3f' % (d['ms_per_layer_call'] * 1e3, d['frac'], d['graph_ms_per_layer_call'] * 1e3, d['graph_frac']))
PY
