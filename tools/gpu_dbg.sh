for i in 1 2 3; do echo "--- PDL on run $i"; timeout 100 python tools/ipa_debug.py 8,350 2,37 2>&1 | grep -v Warning | tail -4; done
echo "--- PDL on, launch blocking"; CUDA_LAUNCH_BLOCKING=1 timeout 100 python tools/ipa_debug.py 8,350 2>&1 | grep -v Warning | tail -3
