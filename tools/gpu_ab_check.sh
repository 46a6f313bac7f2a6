# A/B of an environment knob on the training-step bench: bash tools/gpu_ab_check.sh VAR v1 v2 ...
VAR=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  for rep in 1 2; do
    env $VAR=$v timeout 600 python bench.py --no-cpu-baseline --steps 40 2> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$VAR=$v', 'ms_per_step %.4f' % d['ms_per_step'], 'fwd %.4f' % d['forward']['ms_per_step'])"
  done
done
