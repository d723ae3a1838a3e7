export ABX_IPA_RESCALE_GAP=1.0
for R in 17 20; do echo "--- debug 2,350 gap 1 R=$R"; ABX_IPA_ROWS=$R timeout 150 python tools/ipa_debug.py 2,350 2>&1 | grep -v Warning | tail -2 | cut -c1-300; done
unset ABX_IPA_RESCALE_GAP
run() { echo "--- stress $*"; timeout 150 python tools/ipa_stress.py $* 2>&1 | grep -v "Warning\|^$\|Search for\|might be\|For debugging\|Compile with\|File \|\^\^\^" | tail -3; }
run 8 350 16 2.0 0.1
run 8 350 8 8.0 0.1
echo "--- sizes"; timeout 150 python tools/ipa_debug.py 1,1 1,7 2,37 5,131 2>&1 | grep -v Warning | grep "B=" | cut -c1-120
