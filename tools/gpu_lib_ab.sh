# A/B of two builds of the library on one box: bash tools/gpu_lib_ab.sh path/to/other.so   (default build vs the other, alternating)
OTHER=$1
mkdir -p gpurun_out
run() {
  env $1 timeout 300 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 40 2>> gpurun_out/ab_env.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  ms_per_step %.4f  fwd %.4f  e2e %.4f' % (d['ms_per_step'], d['forward']['ms_per_step'], d['e2e']['ms_per_step']))"
}
for i in 1 2 3; do
  echo "[default build]"; run "B200_X=0"
  echo "[$OTHER]"; run "B200_LIB_PATH=$PWD/$OTHER"
done
