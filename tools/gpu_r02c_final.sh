# Round-2c evidence run (one GPU): GPU tests, smoke, bench (both arms), launch list, ncu captures, kernel table, profile.
mkdir -p gpurun_out
P=r02c
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${P}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${P}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${P}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${P}_smoke.log
timeout 900 python bench.py > gpurun_out/${P}_bench_n1.json 2> gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_reference_arm.json 2> gpurun_out/bench_ref.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${P}_launches_train_step_b2_128.csv python tools/step_for_ncu.py 2 128 > gpurun_out/ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/${P}_launches_train_step_b2_128.csv 45 > gpurun_out/${P}_launch_list_train_step_b2_128.txt
# the time-dominant kernel of the replayed step (bench `roofline`): DRAM traffic per launch at level 0
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_bwd_reduce2 --launch-skip 1 -c 1 -f -o gpurun_out/${P}_prof_gn_bwd_reduce python tools/gn_for_ncu.py 2 128 16 > gpurun_out/ncu_gn.log 2>&1
python tools/ncu_kernel_summary.py gpurun_out/${P}_prof_gn_bwd_reduce.ncu-rep "gn_bwd_reduce2 C=16 @ 2x128^3 (alone)" > gpurun_out/${P}_ncu_gn_bwd_reduce_summary.txt 2>&1
timeout 300 python tools/step_profile.py 2 128 5 > gpurun_out/${P}_step_profile_graph_replay.txt 2>&1
timeout 300 python tools/perf_probe.py 2 128 > gpurun_out/${P}_kernel_table.txt 2>&1
timeout 300 python tools/step_timeline.py 2 128 gpurun_out/${P}_step_timeline.csv > /dev/null 2>&1
tail -3 gpurun_out/${P}_pytest_gpu.log; tail -2 gpurun_out/${P}_smoke.log; python -c "
import json; d=json.loads(open('gpurun_out/${P}_bench_n1.json').read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline'], d.get('gpu_baseline'), d['cpu_baseline'], d['clocks'])"
head -12 gpurun_out/${P}_launch_list_train_step_b2_128.txt; cat gpurun_out/${P}_ncu_gn_bwd_reduce_summary.txt | head -12
