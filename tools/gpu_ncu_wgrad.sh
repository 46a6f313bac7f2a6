mkdir -p gpurun_out
python tools/wgrad_profile.py 2 64 32 2>&1 | tail -1
python tools/wgrad_profile.py 2 32 64 2>&1 | tail -1
python tools/wgrad_profile.py 2 16 128 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_gemm --launch-skip 3 -c 1 -f -o gpurun_out/prof_wgrad32 python tools/wgrad_profile.py 2 64 32 > gpurun_out/ncu_wgrad32.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_gemm --launch-skip 3 -c 1 -f -o gpurun_out/prof_wgrad128 python tools/wgrad_profile.py 2 16 128 > gpurun_out/ncu_wgrad128.log 2>&1
ls -la gpurun_out/*.ncu-rep
