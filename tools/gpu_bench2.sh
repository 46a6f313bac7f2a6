mkdir -p gpurun_out
for v in 1 2 3; do
  timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 40 2> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('run $v', 'ms_per_step %.4f' % d['ms_per_step'], 'fwd %.4f' % d['forward']['ms_per_step'], 'e2e %.4f' % d['e2e']['ms_per_step'], 'launches', d['gpu_launches']/d['steps'])" || tail -5 gpurun_out/ab.err
done
env B200_BWD_SCHED=0 B200_D2S_FUSED=0 B200_WGL_NO_HALF=1 timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 40 2> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('session-start switches (old order, d2s pass, both chunks)', 'ms_per_step %.4f' % d['ms_per_step'], 'fwd %.4f' % d['forward']['ms_per_step'])"
timeout 300 python tools/step_timeline.py 2 128 gpurun_out/r02f_step_timeline.csv 2>&1 | tail -1
timeout 900 python -m pytest tests/test_layer_parity_gpu.py tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -4
