#!/bin/bash
mkdir -p gpurun_out
for d in 1 2 3 6; do
  echo "== kb_per_drain $d"
  ABX_GEMM_KB_PER_DRAIN=$d timeout 150 python -m pytest tests/test_gpu_gemm.py -x -q 2>&1 | tail -2
  ABX_GEMM_KB_PER_DRAIN=$d timeout 120 python tools/bench_gemm.py 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    try: r = json.loads(ln)
    except Exception: continue
    print(r['M'], r['N'], r['K'], 'bn128 %.3f' % r['abx_bn128_ms'], 'bn64 %.3f' % r['abx_bn64_ms'], 'torch %.3f' % r['torch_ms'], 'err %.2e torch_err %.2e' % (r['abx_maxerr'], r['torch_maxerr']), 'gate_res %.3f' % r.get('abx_gate_res_ms', 0))
"
done
