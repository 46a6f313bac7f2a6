#!/usr/bin/env python
"""Benchmark of the ResUNet hot path (BASELINE.json metric: voxels/sec fwd & fwd+bwd, 4x128^3).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference arithmetic on host CPU cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...          # one rank per GPU (weak scaling)

One "step" is one training step of the drop-in `UNet` on a batch of synthetic 4x128^3 volumes:
forward + Dice_loss_joint + backward (+ gradient all-reduce for N > 1) + the stock Adam step the
reference trainer runs (train.py:201-221).  `value` = whole-job voxels/s with inputs resident in
HBM; `e2e` = the same step through the public API with pinned HOST inputs (H2D inside the timed
region) and the loss read back to the host every step.  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

FLOP_PER_VOXEL_FWD = 142752
FLOP_PER_VOXEL_FWD_BWD = 424800


def read_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 50 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1]); power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power) if power else None}


def init_weights(model, seed):
    """Same distributions as the reference's weight_init.py:22-27 (kaiming normal a=1e-2 for Conv3d
    weights, N(0,1) conv bias, GroupNorm untouched); own generator so every rank draws the same."""
    import math
    import torch
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.Conv3d):
            fan_in = m.weight.shape[1] * m.weight[0, 0].numel()
            std = math.sqrt(2.0 / (1 + 1e-2 ** 2)) / math.sqrt(fan_in)
            m.weight.data.copy_(torch.randn(m.weight.shape, generator=g) * std)
            if m.bias is not None:
                m.bias.data.copy_(torch.randn(m.bias.shape, generator=g))


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the oracle restatement of the reference arithmetic on host cores
# ------------------------------------------------------------------------------------------------
def cpu_train_step_seconds(size, reps, batch=1):
    import torch
    from oracle import resunet_oracle as O      # the checker, timed as the CPU baseline only
    sd = O.init_params(1337)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(batch, 4, size, size, size, generator=g)
    t = (torch.rand(batch, 3, size, size, size, generator=g) > 0.7).float()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        O.train_step(sd, x, t)
        best = min(best, time.perf_counter() - t0)
    return best


def run_reference(args, rank):
    if rank != 0:
        return
    import torch
    # torchrun exports OMP_NUM_THREADS=1 for every rank; the reference arm is rank 0 alone on the host, so it takes
    # every core this process may run on ("all the host threads it can use")
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    if torch.get_num_threads() < avail:
        torch.set_num_threads(avail)
    cores = torch.get_num_threads()
    steps, warm = args.steps, args.warmup
    # bounded sample: pick the largest cube whose (K+W) steps fit ~150 s, from a 32^3 calibration
    t32 = cpu_train_step_seconds(32, 1)
    t32 = min(t32, cpu_train_step_seconds(32, 1))
    budget = 150.0 / max(1, steps + warm)
    size = 32
    for s in (128, 96, 64, 48):
        if t32 * (s / 32.0) ** 3 * 1.3 <= budget:
            size = s
            break
    from oracle import resunet_oracle as O
    sd = O.init_params(1337)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, size, size, size, generator=g)
    t = (torch.rand(1, 3, size, size, size, generator=g) > 0.7).float()
    for _ in range(warm):
        O.train_step(sd, x, t)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(sd, x, t)
    dt = (time.perf_counter() - t0) / max(1, steps)
    v = size ** 3 / dt
    sample = "1 x 4x%d^3 crop per step: fwd + Dice_loss_joint + autograd backward, fp32, torch CPU (oneDNN)" % size
    emit({
        "impl": "reference", "metric": "voxels/sec fwd+bwd (ResUNet train step, 4x128^3 volumes)", "value": v,
        "unit": "voxels/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "train step fwd+Dice+bwd, 4x128^3 volumes (bounded sample: %s)" % sample},
        "cpu_baseline": {"value": v, "unit": "voxels/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle port of /root/reference model.py+loss.py (reference is Python and does not travel to the "
                "GPU box); excludes the reference's gc.collect() per forward and the optimizer step",
    })


# ------------------------------------------------------------------------------------------------
_JSON_FD = None


def claim_stdout():
    """The contract is ONE JSON line on stdout.  NCCL prints its version banner to fd 1 when the communicator is
    created, and the reference-compatible UNet constructor prints 'UNet [...]' (model.py:311): everything except
    the result line goes to stderr; the result is written to the saved descriptor."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, line)
    else:
        os.write(_JSON_FD, line)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=2, help="volumes per GPU per step (config 3: 2)")
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels one by one instead of one CUDA graph per step")
    ap.add_argument("--torch-adam", action="store_true", help="stock torch.optim.Adam(fused=True) instead of brats2019_b200.optim.FusedAdam")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if args.warmup < 3:
        args.warmup = 3

    import torch
    import torch.distributed as dist
    import brats2019_b200 as B
    from brats2019_b200 import ops
    from brats2019_b200.parallel import DistributedUNet

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # Everything below runs on a non-default stream: work queued on the legacy default stream
    # implicitly synchronises with every blocking stream and breaks CUDA-graph capture.
    main_stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(main_stream):
        run(args, rank, world, local_rank, dev)
    if world > 1:
        dist.destroy_process_group()


def run(args, rank, world, local_rank, dev):
    import torch
    import torch.distributed as dist
    import brats2019_b200 as B
    from brats2019_b200 import ops
    from brats2019_b200.parallel import DistributedUNet
    peaks = read_peaks()
    Bsz, S = args.batch, args.size
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):      # the reference-compatible ctor prints 'UNet [...]' (model.py:311)
        model = B.UNet(**B.DEFAULT_CFG)
    init_weights(model, 1337)
    model = model.to(dev).train()
    crit = B.Dice_loss_joint(index=0, priority=1)
    net = model
    if world > 1:
        net = DistributedUNet(model)
        crit.process_group = net.process_group
    use_graph = not args.no_graph
    if args.torch_adam:
        opt = torch.optim.Adam(model.parameters(), lr=2e-5, weight_decay=1e-6, amsgrad=True,
                               capturable=use_graph, fused=True)                            # main.py:133-140 (stock torch Adam, fused impl)
    else:
        from brats2019_b200.optim import FusedAdam
        opt = FusedAdam(model.parameters(), lr=2e-5, weight_decay=1e-6, amsgrad=True, lr_step_size=16000, lr_gamma=0.5,
                        model=net)                                                          # main.py:133-142 as one kernel

    g = torch.Generator().manual_seed(100 + rank)
    x_host = torch.randn(Bsz, 4, S, S, S, generator=g).pin_memory()
    t_host = (torch.rand(Bsz, 3, S, S, S, generator=g) > 0.7).float().pin_memory()
    x_dev, t_dev = x_host.to(dev), t_host.to(dev)

    def step_resident():
        opt.zero_grad(set_to_none=True)
        loss = crit(net([x_dev]), [t_dev])
        loss.backward()
        opt.step()
        return loss

    def step_e2e():
        xd = x_host.to(dev, non_blocking=True)
        td = t_host.to(dev, non_blocking=True)
        opt.zero_grad(set_to_none=True)
        loss = crit(net([xd]), [td])
        loss.backward()
        opt.step()
        return loss.item()                       # device -> host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        l0 = ops.LAUNCHES[0]
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), ops.LAUNCHES[0] - l0

    for _ in range(args.warmup):
        step_resident()
    if use_graph:
        # the whole step as ONE CUDA-graph launch (brats2019_b200.graphs): same kernels, same work,
        # no per-launch Python/driver overhead.  Static input buffers are refreshed by copy_.
        from brats2019_b200.graphs import GraphedForward, GraphedTrainStep
        gstep = GraphedTrainStep(net, crit, opt, x_dev, t_dev, warmup=1)
        launches_per_step = None

        def step_resident():          # noqa: F811
            return gstep()

        gstep.prefetch(x_host, t_host)               # prime the input pipeline (outside the timed region)

        def step_e2e():               # noqa: F811
            # every step: one H2D copy of a full host batch (issued on the copy stream so that it
            # overlaps the previous step's kernels), the graph, and a D2H read of the loss
            loss = gstep.step_prefetched()
            gstep.prefetch(x_host, t_host)
            return loss.item()
        l0 = ops.LAUNCHES[0]
        GraphedTrainStep._eager(gstep)                # count kernels of one eager step (same as the graph's)
        launches_per_step = ops.LAUNCHES[0] - l0
        for _ in range(2):
            step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_total, launches = timed(step_resident, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    if use_graph:
        launches = launches_per_step * args.steps
    ms_step = ms_total / args.steps
    vox_step = Bsz * S ** 3 * world
    value = vox_step / (ms_step * 1e-3)

    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    e2e_value = vox_step / (ms_e2e / args.steps * 1e-3)

    # forward-only (eval, no_grad) on the same resident batch
    model.eval()
    with torch.no_grad():
        for _ in range(3):
            net([x_dev])
        if use_graph:
            gf = GraphedForward(net, x_dev)
            ms_fwd, _ = timed(lambda: gf(), args.steps)
        else:
            ms_fwd, _ = timed(lambda: net([x_dev]), args.steps)
    model.train()
    fwd_value = vox_step / (ms_fwd / args.steps * 1e-3)

    # dominant conv kernel, timed alone: level-0 3x3x3 16->16 implicit GEMM (5 per forward, 5 data gradients per
    # backward; 44 % of the model's FLOPs)
    roof = None
    if rank == 0:
        desc = ops.conv_desc(ops.MODE_K3, Bsz, S, S, S, 16, 16)
        xa = ops.act_zeros(Bsz, S, S, S, 16, dev)
        xa.interior().copy_(torch.randn(2, Bsz, S, S, S, 8, device=dev).to(torch.bfloat16))
        w = torch.randn(16, 16, 3, 3, 3, device=dev) * 0.05
        pk = ops.conv_pack_weight(desc, ops.W_FWD, w)
        out = ops.act_zeros(Bsz, S, S, S, 16, dev)
        st = torch.empty(ops.conv_ctas(desc) * Bsz * 16, device=dev)
        for _ in range(3):
            ops.conv_run(desc, xa, pk, out, stats=st)
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            ops.conv_run(desc, xa, pk, out, stats=st)
        e1.record()
        torch.cuda.synchronize()
        k_ms = e0.elapsed_time(e1) / reps
        flops = 2.0 * Bsz * S ** 3 * 16 * 432
        ach = flops / (k_ms * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(REPO, "profiles", "dominant_kernel_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roof = {"bound": "tensor", "achieved": ach, "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                "frac": ach / peaks["tf_burst"], "traffic": traffic,
                "kernel": "conv_band_kernel<BF16> (3x3x3 16->16 implicit GEMM, kd+kh folded, +GN stats) @ %dx%d^3" % (Bsz, S),
                "ms_per_launch": k_ms,
                "algorithmic_flops_per_launch": flops, "peak_source": peaks["source"] + ", burst (kernel timed alone)"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import torch as _t
        secs = cpu_train_step_seconds(128, 2)
        cpu = {"value": 128 ** 3 / secs, "unit": "voxels/s", "cores": _t.get_num_threads(), "kind": "port",
               "sample": "1 x 4x128^3 train step (fwd + Dice + autograd bwd), fp32 torch CPU, best of 2, %.1f s" % secs}

    if rank == 0:
        flops_step = FLOP_PER_VOXEL_FWD_BWD * vox_step
        out = {
            "metric": "voxels/sec fwd+bwd (ResUNet train step, 4x128^3 volumes)", "value": value, "unit": "voxels/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "train step: fwd + Dice_loss_joint + bwd%s + Adam(amsgrad), batch %d x 4x%d^3 per GPU "
                                   "(BASELINE config 3%s)" % (" + NCCL grad all-reduce" if world > 1 else "", Bsz, S,
                                                              "/4" if world > 1 else ""),
                       "per_gpu_batch": Bsz, "global_batch": Bsz * world, "volume": [4, S, S, S],
                       "parallelism": "dp%d" % world, "cuda_graph": bool(use_graph),
                       "l2": "no flush needed: per-step working set (>2 GB of activations) >> 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": "voxels/s", "h2d_bytes_per_step": int(x_host.numel() * 4 + t_host.numel() * 4),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "forward": {"value": fwd_value, "unit": "voxels/s", "ms_per_step": ms_fwd / args.steps,
                        "tensor_frac_of_sustained": fwd_value / world * FLOP_PER_VOXEL_FWD / 1e12 / peaks["tf_sust"]},
            "model_tflops": flops_step / (ms_step * 1e-3) / 1e12 / world,
            "model_tensor_frac_of_sustained": flops_step / (ms_step * 1e-3) / 1e12 / world / peaks["tf_sust"],
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
        }
        emit(out)


if __name__ == "__main__":
    main()
