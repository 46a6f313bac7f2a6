mkdir -p gpurun_out
timeout 300 python tests/gpu_opcheck.py ew > gpurun_out/opcheck_ew.log 2>&1; echo "opcheck rc=$?"
grep -c PASS gpurun_out/opcheck_ew.log; grep -v PASS gpurun_out/opcheck_ew.log | head -20
timeout 300 python tools/perf_probe.py 2 128 2>&1 | grep -E "upsample|pack_input|sigmoid"
bash tools/gpu_quick.sh 2
