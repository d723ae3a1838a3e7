#!/bin/bash
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_trunk_ops.py -k layernorm tests/test_gpu_model.py::test_trunk_and_score_network_match_reference tests/test_gpu_model.py::test_ipascore_matches_reference -x -q > gpurun_out/ln_check.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ln_check.log
tail -3 gpurun_out/ln_check.log
timeout 30 python tools/ln_one.py > gpurun_out/ln_time.log 2>&1; cat gpurun_out/ln_time.log | tail -2
