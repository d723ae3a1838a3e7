#!/bin/bash
# A/B of programmatic dependent launch and the L2 warm-up inside the IPA layer-call (B200, N=350)
mkdir -p gpurun_out
OUT=gpurun_out/ipa_pdl_ab.jsonl; : > $OUT
timeout 300 python -m pytest tests/test_gpu_ipa.py -x -q 2>&1 | tail -3
ABX_IPA_PREFETCH_MB=32 timeout 300 python -m pytest tests/test_gpu_ipa.py -x -q 2>&1 | tail -2
for B in 8 4; do
  for cfg in "0 0" "1 0" "1 16" "1 32" "1 64" "1 96"; do
    set -- $cfg
    ABX_IPA_PDL=$1 ABX_IPA_PREFETCH_MB=$2 timeout 120 python tools/bench_ipa.py --B $B --N 350 --iters 30 --graph 8 >> $OUT 2>gpurun_out/ipa_pdl_ab.err || tail -3 gpurun_out/ipa_pdl_ab.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/ipa_pdl_ab.jsonl'):
    d = json.loads(l)
    print(d['B'], 'pdl', d['pdl'], 'pf', d['prefetch_mb'], 'eager us %.1f frac %.3f | graph us %.1f frac %.3f' % (d['ms_per_layer_call'] * 1e3, d['frac'], d['graph_ms_per_layer_call'] * 1e3, d['graph_frac']))
PY
timeout 400 python -m pytest tests/test_gpu_model.py -x -q 2>&1 | tail -3
