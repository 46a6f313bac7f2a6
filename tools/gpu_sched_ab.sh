# A/B of the backward launch order (B200_BWD_SCHED) + timeline of the new order + the parity tests that cover it
mkdir -p gpurun_out
for v in 0 1 0 1; do
  env B200_BWD_SCHED=$v timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 40 2> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('B200_BWD_SCHED=$v', 'ms_per_step %.4f' % d['ms_per_step'], 'fwd %.4f' % d['forward']['ms_per_step'], 'e2e %.4f' % d['e2e']['ms_per_step'])"
done
timeout 300 python tools/step_timeline.py 2 128 gpurun_out/r02b_step_timeline.csv 2>&1 | tail -1
timeout 900 python -m pytest tests/test_layer_parity_gpu.py tests/test_model_gpu.py tests/test_optim_gpu.py -m gpu -x -q 2>&1 | tail -4
