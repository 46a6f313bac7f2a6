# A/B: what runs where at the start of a step (B200_FWD_PACK = 1 default / 0 serial / 2 swapped), bench + timeline
mkdir -p gpurun_out
for m in 1 0 2 1; do
  echo "[B200_FWD_PACK=$m]"
  for i in 1 2; do
    B200_FWD_PACK=$m timeout 300 python bench.py --no-cpu-baseline --no-gpu-baseline --steps 40 2>> gpurun_out/ab_env.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  ms_per_step %.4f  fwd %.4f  e2e %.4f' % (d['ms_per_step'], d['forward']['ms_per_step'], d['e2e']['ms_per_step']))"
  done
  B200_FWD_PACK=$m timeout 200 python tools/step_timeline.py 2 128 gpurun_out/tl_fwdpack_$m.csv --all > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.DictReader(open('gpurun_out/tl_fwdpack_$m.csv')))
idx=[i for i,r in enumerate(rows) if 'adam' in r['name']]
i=idx[0]
t0=float(rows[i]['start_us'])+float(rows[i]['dur_us'])
for r in rows[i+1:i+5]:
    print('   +%.1f us  %.1f  %s  s%s' % (float(r['start_us'])-t0, float(r['dur_us']), r['name'][:34], r['stream']))
print('   period %.1f us' % (float(rows[idx[1]]['start_us'])-float(rows[idx[0]]['start_us'])))
PY
done
