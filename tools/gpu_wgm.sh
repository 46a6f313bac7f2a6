mkdir -p gpurun_out
timeout 180 python tests/gpu_opcheck.py wgrad > gpurun_out/opcheck_wgrad.log 2>&1; echo "opcheck rc=$?"
grep -c PASS gpurun_out/opcheck_wgrad.log; grep -v PASS gpurun_out/opcheck_wgrad.log | head -30
for cfg in "2 128 16" "2 64 32" "2 32 64" "2 16 128"; do timeout 120 python tests/wgrad_profile.py $cfg 2>&1 | tail -1; done
bash tools/gpu_quick.sh 2
