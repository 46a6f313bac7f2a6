for v in 0 1 0 1; do
  env B200_PDL_SMALL=$v timeout 600 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 40 2> gpurun_out/ab.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('B200_PDL_SMALL=$v', 'ms_per_step %.4f' % d['ms_per_step'], 'fwd %.4f' % d['forward']['ms_per_step'], 'e2e %.4f' % d['e2e']['ms_per_step'])" || tail -5 gpurun_out/ab.err
done
