mkdir -p gpurun_out
timeout 180 python tests/gpu_opcheck.py wgrad > gpurun_out/opcheck_wgrad.log 2>&1; echo "opcheck rc=$?"
grep -c PASS gpurun_out/opcheck_wgrad.log; grep -v PASS gpurun_out/opcheck_wgrad.log | head -30
timeout 120 python tests/wgrad_profile.py 2 128 16 2>&1 | tail -2
B200_NO_WGRAD_MARCH=1 timeout 120 python tests/wgrad_profile.py 2 128 16 2>&1 | tail -1
