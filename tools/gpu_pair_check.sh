# Pair-mode line weight gradient: op check (all forms), A/B timing alone, plan sweep.
mkdir -p gpurun_out
timeout 600 python tests/gpu_opcheck.py wgrad > gpurun_out/pair_opcheck.log 2>&1; echo "opcheck rc=$?"
grep -c OK gpurun_out/pair_opcheck.log; grep -v OK gpurun_out/pair_opcheck.log | tail -15
timeout 600 python tools/wgrad_ab.py 2 128 --pair-sweep --contend 2>&1 | tee gpurun_out/pair_ab.txt
