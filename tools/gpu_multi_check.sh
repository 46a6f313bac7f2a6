# usage (on a box with N GPUs): bash tools/gpu_multi_check.sh N
N=${1:-2}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR tools/ddp_parity.py > gpurun_out/ddp_parity_n$N.log 2>&1; echo "rc=$?" >> gpurun_out/ddp_parity_n$N.log
timeout 600 $TR bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?" >> gpurun_out/bench_n$N.err
timeout 900 python tools/bench_inference.py > gpurun_out/inference_n1.jsonl 2> gpurun_out/inference_n1.err; echo "rc=$?" >> gpurun_out/inference_n1.err
timeout 900 $TR tools/bench_inference.py > gpurun_out/inference_n$N.jsonl 2> gpurun_out/inference_n$N.err; echo "rc=$?" >> gpurun_out/inference_n$N.err
tail -5 gpurun_out/ddp_parity_n$N.log; cat gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/inference_n1.jsonl; tail -5 gpurun_out/inference_n1.err; cat gpurun_out/inference_n$N.jsonl
