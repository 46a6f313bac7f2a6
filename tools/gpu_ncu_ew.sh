mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"upsample2x|pack_input|sigmoid_bwd_pack2|d2s_kernel" -f -o gpurun_out/prof_ew python tools/step_for_ncu.py 2 128 > gpurun_out/ncu_ew.log 2>&1
ls -la gpurun_out/prof_ew.ncu-rep; tail -3 gpurun_out/ncu_ew.log
