# A/B of the per-kernel launch priorities (B200_PRIO="conv wgrad") on the training-step bench
mkdir -p gpurun_out
for v in "0 0" "-2 -1" "0 -1" "-1 -1" "-1 -2" "0 0" "-2 -1"; do
  env B200_PRIO="$v" timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 40 2> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('B200_PRIO=$v', 'ms_per_step %.4f' % d['ms_per_step'], 'fwd %.4f' % d['forward']['ms_per_step'], 'e2e %.4f' % d['e2e']['ms_per_step'])"
done
timeout 300 python tools/step_timeline.py 2 128 gpurun_out/r02c_step_timeline.csv 2>&1 | tail -1
