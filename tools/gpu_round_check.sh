set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 300 python tools/step_profile.py 2 128 5 > gpurun_out/step_profile.txt 2>&1
timeout 300 python tools/perf_probe.py 2 128 > gpurun_out/perf_probe.txt 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/step_for_ncu.py 2 128 > gpurun_out/ncu_step.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_band --launch-skip 3 -c 1 -f -o gpurun_out/prof_conv_band python tools/conv_for_ncu.py 2 128 16 > gpurun_out/ncu_conv.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench_n1.json
