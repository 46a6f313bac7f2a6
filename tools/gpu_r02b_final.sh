# Round-2b evidence run (one GPU): GPU tests, smoke, bench (both arms), launch list, ncu captures, kernel table, profile.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02b_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02b_smoke.log
timeout 900 python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_reference_arm.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02b_launches.csv python tools/step_for_ncu.py 2 128 > gpurun_out/ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/r02b_launches.csv 45 > gpurun_out/r02b_launch_list_train_step_b2_128.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_line_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r02b_prof_wgrad_line python tools/wgrad_profile.py 2 128 16 > gpurun_out/ncu_wl.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_gemm_kernel --launch-skip 3 -c 1 -f -o gpurun_out/r02b_prof_wgrad_gemm_c32 python tools/wgrad_profile.py 2 64 32 > gpurun_out/ncu_wg.log 2>&1
timeout 300 python tools/step_profile.py 2 128 5 > gpurun_out/r02b_step_profile_graph_replay.txt 2>&1
timeout 300 python tools/perf_probe.py 2 128 > gpurun_out/r02b_kernel_table.txt 2>&1
timeout 300 python tools/step_timeline.py 2 128 gpurun_out/r02b_final_step_timeline.csv > /dev/null 2>&1
tail -3 gpurun_out/r02b_pytest_gpu.log; tail -2 gpurun_out/r02b_smoke.log; python -c "
import json; d=json.loads(open('gpurun_out/r02b_bench_n1.json').read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline'], d.get('gpu_baseline'), d['cpu_baseline'])"
head -12 gpurun_out/r02b_launch_list_train_step_b2_128.txt
