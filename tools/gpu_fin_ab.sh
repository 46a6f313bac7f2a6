# last-CTA finalize (forward statistics, GroupNorm-backward sums): op checks, A/B on the training step, parity tests
mkdir -p gpurun_out
timeout 600 python tests/gpu_opcheck.py ew conv3 2>&1 | grep -v "^PASS" | tail -30
for v in "0 0" "1 1" "1 0" "0 1" "0 0" "1 1"; do
  set -- $v
  env B200_GN_FUSED_STATS=$1 B200_GN_BWD_FUSED_FIN=$2 timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 40 2> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('stats=$1 bwdfin=$2', 'ms_per_step %.4f' % d['ms_per_step'], 'fwd %.4f' % d['forward']['ms_per_step'], 'e2e %.4f' % d['e2e']['ms_per_step'], 'launches', d['gpu_launches']/d['steps'])" || tail -5 gpurun_out/ab.err
done
timeout 900 python -m pytest tests/test_layer_parity_gpu.py tests/test_model_gpu.py tests/test_optim_gpu.py -m gpu -x -q 2>&1 | tail -4
