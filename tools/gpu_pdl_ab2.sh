mkdir -p gpurun_out
for v in 1 2 1 2 0; do
  env B200_PDL=$v timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 40 2> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('B200_PDL=$v', 'ms_per_step %.4f' % d['ms_per_step'], 'fwd %.4f' % d['forward']['ms_per_step'], 'e2e %.4f' % d['e2e']['ms_per_step'])" || tail -5 gpurun_out/ab.err
done
timeout 800 python tests/gpu_opcheck.py ew conv3 conv1 2>&1 | grep -v "^PASS" | tail -8
timeout 900 python -m pytest tests/test_layer_parity_gpu.py tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -3
