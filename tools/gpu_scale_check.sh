# bench.py at N GPUs exactly as the driver launches it: bash tools/gpu_scale_check.sh N
N=${1:-4}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; wc -l gpurun_out/bench_n$N.json; cut -c1-300 gpurun_out/bench_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
echo "ref rc=$?"; wc -l gpurun_out/bench_ref_n$N.json; cut -c1-200 gpurun_out/bench_ref_n$N.json
