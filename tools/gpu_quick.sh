# GPU regression + bench in one call: bash tools/gpu_quick.sh [bench repeats]
R=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for i in $(seq $R); do
  timeout 600 python bench.py --no-cpu-baseline --steps 40 2> gpurun_out/quick.err | tee gpurun_out/bench_quick_$i.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('ms_per_step %.4f' % d['ms_per_step'], 'fwd %.4f' % d['forward']['ms_per_step'], 'e2e %.4f' % d['e2e']['ms_per_step'], 'launches/step', d['gpu_launches']/d['steps'])"
done
