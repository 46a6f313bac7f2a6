for v in 1 2 3; do
  timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 40 2> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('run $v', 'ms_per_step %.4f' % d['ms_per_step'], 'fwd %.4f' % d['forward']['ms_per_step'], 'e2e %.4f' % d['e2e']['ms_per_step'])" || tail -5 gpurun_out/ab.err
done
env B200_NO_WGRAD_OVERLAP=1 timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 40 2> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no side stream', 'ms_per_step %.4f' % d['ms_per_step'], 'fwd %.4f' % d['forward']['ms_per_step'])"
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_launch.py tests/test_staging_gpu.py -m gpu -x -q 2>&1 | tail -3
