timeout 400 python -m pytest tests/test_gpu_ipa.py -q 2>&1 | tail -4
for B in 8 4 1; do timeout 120 python tools/bench_ipa.py --B $B --N 350 --profile 1 --graph 8 > gpurun_out/bench_ipa_fused_B${B}_r02e.log 2>&1; grep -v "^ *0.0 us\|arn" gpurun_out/bench_ipa_fused_B${B}_r02e.log | tail -6 | cut -c1-600; done
timeout 120 python tools/bench_ipa.py --B 8 --N 262 --graph 8 > gpurun_out/bench_ipa_fused_B8_N262_r02e.log 2>&1; tail -1 gpurun_out/bench_ipa_fused_B8_N262_r02e.log | cut -c1-600
