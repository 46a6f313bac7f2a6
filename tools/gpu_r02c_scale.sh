# Round-2c multi-GPU evidence on N GPUs of one box (short: weak-scaling line + config 4 as written): bash tools/gpu_r02b_scale.sh N
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29533 bench.py --gpus $N --steps 40 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02c_bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "weak rc=$?"
timeout 500 $TR --master-port 29534 bench.py --gpus $N --global-batch 16 --steps 40 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/r02c_bench_n${N}_global16.json 2> gpurun_out/bench_g16_n$N.err; echo "global16 rc=$?"
python - <<PY
import json
for f in ("gpurun_out/r02c_bench_n$N.json", "gpurun_out/r02c_bench_n${N}_global16.json"):
    try:
        d = json.loads(open(f).read())
        print(f, {k: d[k] for k in ("value", "ms_per_step", "n_gpus", "scaling")}, "e2e %.3f ms" % d["e2e"]["ms_per_step"], d.get("ddp_parity"))
    except Exception as e:
        print(f, "unreadable", e)
PY
