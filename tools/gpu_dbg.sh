timeout 150 python tools/ipa_debug.py 1,1 1,7 2,37 3,100 5,131 8,350 2>&1 | grep "B=\|FAILED\|rror" | cut -c1-150
timeout 120 python tools/bench_ipa.py --B 8 --N 350 --profile 1 --prof 1 --graph 8 > gpurun_out/bench_ipa_fused_B8_q30.log 2>&1; grep "loop\|extra\|ipa_fused_kernel\|ms_per_layer\|rror" gpurun_out/bench_ipa_fused_B8_q30.log | cut -c1-230
export ABX_IPA_RESCALE_GAP=1.0
for R in 20 17; do echo "--- debug 2,350 gap 1 R=$R"; ABX_IPA_ROWS=$R timeout 150 python tools/ipa_debug.py 2,350 2>&1 | grep -v Warning | tail -2 | cut -c1-200; done
unset ABX_IPA_RESCALE_GAP
echo "--- stress"; timeout 150 python tools/ipa_stress.py 8 350 8 8.0 0.1 2>&1 | grep -v "Warning\|^$" | tail -2
