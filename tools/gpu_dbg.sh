timeout 300 python -m pytest tests/test_gpu_trunk_ops.py -q -x -k "pair_attention" 2>&1 | tail -3
timeout 120 python tools/bench_attention.py 2>&1 | tail -3
