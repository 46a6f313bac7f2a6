# Round-2 evidence run (one GPU): launch list of one training step, full ncu captures of the two weight-gradient
# kernels, per-kernel table, step profile.  bash tools/gpu_r02_profiles.sh
mkdir -p gpurun_out
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/step_for_ncu.py 2 128 > gpurun_out/ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches.csv 45 > gpurun_out/r02_launch_list_train_step_b2_128.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_line_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r02_prof_wgrad_line python tools/wgrad_profile.py 2 128 16 > gpurun_out/ncu_wl.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_gemm_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r02_prof_wgrad_gemm_c32 python tools/wgrad_profile.py 2 64 32 > gpurun_out/ncu_wg.log 2>&1
timeout 300 python tools/step_profile.py 2 128 5 > gpurun_out/r02_step_profile_graph_replay.txt 2>&1
timeout 300 python tools/perf_probe.py 2 128 > gpurun_out/r02_kernel_table.txt 2>&1
head -12 gpurun_out/r02_launch_list_train_step_b2_128.txt; tail -30 gpurun_out/r02_kernel_table.txt
