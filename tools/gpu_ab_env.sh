# A/B of environment switches on the replayed training step, two runs per setting on one box:
#   bash tools/gpu_ab_env.sh "NAME=VALUE ..." "NAME=VALUE ..." ...      ("" = defaults)
mkdir -p gpurun_out
run() {
  env $1 timeout 300 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 40 2>> gpurun_out/ab_env.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  ms_per_step %.4f  fwd %.4f  e2e %.4f' % (d['ms_per_step'], d['forward']['ms_per_step'], d['e2e']['ms_per_step']))"
}
for setting in "$@"; do
  echo "[$setting]"
  run "$setting"; run "$setting"
done
