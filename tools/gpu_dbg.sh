bash tools/gpu_quick_ipa.sh q24 2>&1 | grep "B=\|loop\|ipa_fused_kernel\|ms_per_layer\|FAILED\|rror" | cut -c1-220
export ABX_IPA_RESCALE_GAP=1.0
for R in 20; do echo "--- debug 2,350 gap 1 R=$R"; ABX_IPA_ROWS=$R timeout 150 python tools/ipa_debug.py 2,350 2>&1 | grep -v Warning | tail -2 | cut -c1-200; done
unset ABX_IPA_RESCALE_GAP
echo "--- stress"; timeout 150 python tools/ipa_stress.py 8 350 6 8.0 0.1 2>&1 | grep -v "Warning\|^$" | tail -2
