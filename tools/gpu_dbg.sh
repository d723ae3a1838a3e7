timeout 150 python tools/ipa_debug.py 2,37 8,350 2>&1 | grep "B=\|FAILED\|rror" | cut -c1-120
timeout 120 python tools/bench_ipa.py --B 8 --N 350 --profile 1 --prof 1 --graph 8 > gpurun_out/bench_ipa_fused_B8_q36.log 2>&1; grep "loop\|extra\|ipa_fused_kernel\|ms_per_layer\|rror" gpurun_out/bench_ipa_fused_B8_q36.log | cut -c1-330
