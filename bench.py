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
                                          "--format=csv,noheader,nounits", "-lms", "20"],
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
class CpuReferenceStep:
    """One training step of the reference arithmetic on host cores: the oracle restatement of model.py + loss.py
    (fp32 torch CPU ops, autograd backward) followed by the optimizer of main.py:133-138 (torch.optim.Adam, amsgrad,
    weight decay) - the same work the CUDA arm's step does.  The checker, timed as a baseline only."""

    def __init__(self, batch, size):
        import torch
        from oracle import resunet_oracle as O
        self.O, self.torch = O, torch
        self.sd = O.init_params(1337)
        dead = set(O.dead_param_names())
        self.leaves = {k: v.clone().requires_grad_(True) for k, v in self.sd.items() if k not in dead}
        self.opt = torch.optim.Adam(list(self.leaves.values()), lr=2e-5, weight_decay=1e-6, amsgrad=True)
        g = torch.Generator().manual_seed(0)
        self.x = torch.randn(batch, 4, size, size, size, generator=g)
        self.t = (torch.rand(batch, 3, size, size, size, generator=g) > 0.7).float()

    def __call__(self):
        self.opt.zero_grad(set_to_none=True)
        loss = self.O.dice_loss_joint(self.O.unet_forward(self.leaves, [self.x]), [self.t])
        loss.backward()
        self.opt.step()
        return float(loss.detach())


def cpu_train_step_seconds(size, reps, batch=1):
    step = CpuReferenceStep(batch, size)
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        step()
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
    # The CUDA arm's workload is `--batch` volumes of 4 x `--size`^3 per step.  Run exactly that when the whole
    # (K + W)-step run fits ~4 minutes on this host (calibrated on a 32^3 crop); otherwise one volume per step.
    t32 = min(cpu_train_step_seconds(32, 1), cpu_train_step_seconds(32, 1))
    per_volume = t32 * (args.size / 32.0) ** 3 * 1.3
    batch = args.batch if per_volume * args.batch * (steps + warm) <= 240.0 else 1
    size = args.size
    step = CpuReferenceStep(batch, size)
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(1, steps)
    v = batch * size ** 3 / dt
    sample = ("%d x 4x%d^3 volumes per step: fwd + Dice_loss_joint + autograd backward + Adam(amsgrad) step, fp32, torch CPU "
              "(oneDNN), %d threads" % (batch, size, cores))
    emit({
        "impl": "reference", "metric": "voxels/sec fwd+bwd (ResUNet train step, 4x128^3 volumes)", "value": v,
        "unit": "voxels/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "train step: fwd + Dice_loss_joint + bwd + Adam(amsgrad), batch %d x 4x%d^3 per step "
                               "(BASELINE config 3%s)" % (batch, size, "" if batch == args.batch else "; one volume per step: bounded sample"),
                   "per_gpu_batch": batch, "volume": [4, size, size, size]},
        "cpu_baseline": {"value": v, "unit": "voxels/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle port of /root/reference model.py + loss.py (the reference is Python and does not travel to the GPU "
                "box); excludes only the reference's gc.collect() per forward (model.py:409, ~0.1 s, no numerical effect)",
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
    ap.add_argument("--global-batch", type=int, default=0,
                    help="BASELINE config 4 as written: this many volumes per step over ALL ranks (strong scaling; "
                         "per-GPU batch = global / N).  Default 0: --batch volumes per GPU (weak scaling)")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the stock PyTorch/cuDNN bf16-autocast line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels one by one instead of one CUDA graph per step")
    ap.add_argument("--torch-adam", action="store_true", help="stock torch.optim.Adam(fused=True) instead of brats2019_b200.optim.FusedAdam")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.global_batch:
        if args.global_batch % max(1, world):
            raise SystemExit("--global-batch %d is not divisible by %d ranks" % (args.global_batch, world))
        args.batch = args.global_batch // max(1, world)
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

        def step_e2e_fp32():
            # every step: one H2D copy of a full host batch (issued on the copy stream so that it
            # overlaps the previous step's kernels), the graph, and a D2H read of the loss
            loss = gstep.step_prefetched()
            gstep.prefetch(x_host, t_host)
            return loss.item()

        # The same step fed from COMPACT host staging buffers: bf16 images and uint8 binary targets (46 MB instead of
        # 117 MB per step); the packing kernel and the loss kernels convert in registers (b200_pack_input_t,
        # b200_dice_*_t), so the arithmetic is bit-identical to fp32 inputs (the first kernel rounds to bf16 anyway).
        xc_host = x_host.to(torch.bfloat16).pin_memory()
        tc_host = t_host.to(torch.uint8).pin_memory()
        gstep_c = GraphedTrainStep(net, crit, opt, xc_host.to(dev), tc_host.to(dev), warmup=1)
        gstep_c.prefetch(xc_host, tc_host)

        def step_e2e():               # noqa: F811
            loss = gstep_c.step_prefetched()
            gstep_c.prefetch(xc_host, tc_host)
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
    e2e_bytes = int(x_host.numel() * 4 + t_host.numel() * 4)
    e2e_fp32 = None
    if use_graph:
        e2e_bytes = int(xc_host.numel() * 2 + tc_host.numel())
        for _ in range(2):
            step_e2e_fp32()
        ms_f, _ = timed(step_e2e_fp32, args.steps)
        e2e_fp32 = {"value": vox_step / (ms_f / args.steps * 1e-3), "unit": "voxels/s", "ms_per_step": ms_f / args.steps,
                    "h2d_bytes_per_step": int(x_host.numel() * 4 + t_host.numel() * 4), "d2h_bytes_per_step": 4,
                    "staging": "fp32 images + fp32 targets in pinned host memory (what the reference's loaders produce)"}

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

    # ---- roofline of EVERY kernel family of the step (and the time-dominant one as `roofline`) ----------------------
    # Algorithmic work per family: the ledger of brats2019_b200.ops filled by one eager step (2*M*N*K per conv with
    # padding taps counted, bytes at the logical interface for the memory-bound kernels).
    # Device time per family: CUPTI activity records of graph replays of the same step, taken right after the timed
    # region (CUDA events cannot bracket a kernel inside a graph replay; the event-timed number is `ms_per_step`).
    roof, roof_all = None, None
    if rank == 0 or world > 1:
        ops.LEDGER = []
        if use_graph:
            GraphedTrainStep._eager(gstep)
        else:
            step_resident()
        ledger, ops.LEDGER = ops.LEDGER, None
        torch.cuda.synchronize()
        roof_all = kernel_rooflines(step_resident, ledger, peaks, ms_step)      # every rank replays (collectives inside)
    if rank == 0:
        if roof_all:
            top = roof_all[0]
            traffic, traffic_src = None, None
            tp = os.path.join(REPO, "profiles", "dominant_kernel_traffic.json")
            if os.path.exists(tp):
                try:
                    ent = json.load(open(tp)).get(top["kernel"], {})
                    traffic, traffic_src = ent.get("dram_bytes_per_launch"), ent.get("source")
                except Exception:
                    traffic = None
            roof = {"bound": top["bound"], "achieved": top["achieved"], "peak": top["peak"], "unit": top["unit"],
                    "frac": top["frac"], "traffic": traffic, "kernel": top["kernel"], "launches_per_step": top["launches"],
                    "time_share": top["time_share"], "ms_per_launch": top["ms_per_step"] / max(1, top["launches"]),
                    "algorithmic_work_per_launch": top["work_per_step"] / max(1, top["launches"]),
                    "traffic_source": traffic_src,
                    "peak_source": peaks["source"] + ", sustained (kernel timed inside the step)" if top["bound"] == "tensor"
                    else peaks["source"]}

    # ---- stock PyTorch / cuDNN on the same GPU in the same run: the reference arithmetic
    # (the oracle restatement: F.conv3d, F.group_norm, F.interpolate, autograd) under bf16 autocast + fused Adam ----
    gpu_base = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        gpu_base = torch_gpu_baseline(Bsz, S, dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import torch as _t
        secs = cpu_train_step_seconds(S, 2, batch=1)
        cpu = {"value": S ** 3 / secs, "unit": "voxels/s", "cores": _t.get_num_threads(), "kind": "port",
               "sample": "1 x 4x%d^3 train step (fwd + Dice + autograd bwd + Adam), fp32 torch CPU, best of 2, %.1f s" % (S, secs)}

    # ---- multi-GPU correctness in the same run: N ranks == one process on the concatenated batch ----
    ddp = ddp_parity(model, net, crit, rank, world, dev) if world > 1 else None

    if rank == 0:
        flops_step = FLOP_PER_VOXEL_FWD_BWD * vox_step
        strong = bool(args.global_batch)
        out = {
            "metric": "voxels/sec fwd+bwd (ResUNet train step, 4x128^3 volumes)", "value": value, "unit": "voxels/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": "train step: fwd + Dice_loss_joint + bwd%s + Adam(amsgrad), batch %d x 4x%d^3 per GPU "
                                   "(BASELINE config %s)" % (" + NCCL grad all-reduce" if world > 1 else "", Bsz, S,
                                                             "4: global batch %d" % args.global_batch if strong else
                                                             ("3 per GPU" if world > 1 else "3")),
                       "per_gpu_batch": Bsz, "global_batch": Bsz * world, "volume": [4, S, S, S],
                       "parallelism": "dp%d" % world, "cuda_graph": bool(use_graph),
                       "optimizer": "torch.optim.Adam(fused)" if args.torch_adam else "brats2019_b200.optim.FusedAdam",
                       "l2": "no flush needed: per-step working set (>2 GB of activations) >> 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": "voxels/s", "h2d_bytes_per_step": e2e_bytes,
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps,
                    "staging": ("bf16 images + uint8 targets in pinned host memory, converted inside the first kernels"
                                if use_graph else "fp32 images + fp32 targets in pinned host memory")},
            "e2e_fp32_staging": e2e_fp32,
            "gpu_launches": launches,
            "forward": {"value": fwd_value, "unit": "voxels/s", "ms_per_step": ms_fwd / args.steps,
                        "tensor_frac_of_sustained": fwd_value / world * FLOP_PER_VOXEL_FWD / 1e12 / peaks["tf_sust"]},
            "model_tflops": flops_step / (ms_step * 1e-3) / 1e12 / world,
            "model_tensor_frac_of_sustained": flops_step / (ms_step * 1e-3) / 1e12 / world / peaks["tf_sust"],
            "roofline": roof, "roofline_all": roof_all, "cpu_baseline": cpu, "gpu_baseline": gpu_base, "ddp_parity": ddp,
            "clocks": clocks,
        }
        emit(out)


def kernel_rooflines(step_fn, ledger, peaks, ms_step):
    """[{kernel, launches, ms_per_step, time_share, bound, work_per_step, achieved, peak, unit, frac}], by time."""
    import torch
    from torch.profiler import ProfilerActivity, profile
    reps = 3
    try:
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(reps):
                step_fn()
            torch.cuda.synchronize()
        kern = {}
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                name = ev.name.replace("void ", "").replace("b200::", "")
                us = ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
                c = kern.setdefault(name, [0, 0.0])
                c[0] += 1
                c[1] += us
    except Exception as e:                                   # CUPTI unavailable: no per-kernel table, the headline stands
        sys.stderr.write("kernel_rooflines: profiler failed (%r)\n" % (e,))
        return None
    work = {}
    for name, bound, w in ledger:
        d = work.setdefault(name, [bound, 0.0, 0])
        d[1] += w
        d[2] += 1
    fam = {}
    total_us = sum(v[1] for v in kern.values()) / reps
    for name, (cnt, us) in kern.items():
        key = next((k for k in work if name.startswith(k)), None)
        if key is None:
            key = name.split("(")[0][:60]
        f = fam.setdefault(key, [0, 0.0])
        f[0] += cnt / reps
        f[1] += us / reps
    rows = []
    for key, (cnt, us) in fam.items():
        bound, w, _ = work.get(key, ("hbm", 0.0, 0))
        secs = us * 1e-6
        if bound == "tensor":
            ach, peak, unit = w / secs / 1e12, peaks["tf_sust"], "TFLOP/s"
        else:
            ach, peak, unit = w / secs / 1e9, peaks["hbm"], "GB/s"
        rows.append({"kernel": key, "launches": int(round(cnt)), "ms_per_step": us / 1e3, "time_share": us / total_us,
                     "bound": bound, "work_per_step": w, "achieved": ach, "peak": peak, "unit": unit,
                     "frac": ach / peak if w else None})
    rows.sort(key=lambda r: -r["ms_per_step"])
    return rows[:24]


def torch_gpu_baseline(Bsz, S, dev):
    """The same training step as stock PyTorch ops (cuDNN / ATen) under bf16 autocast on this GPU."""
    import torch
    from oracle import resunet_oracle as O      # the reference arithmetic, timed as a baseline only
    try:
        sd = {k: v.to(dev) for k, v in O.init_params(1337).items()}
        dead = set(O.dead_param_names())
        leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items() if k not in dead}
        opt = torch.optim.Adam(list(leaves.values()), lr=2e-5, weight_decay=1e-6, amsgrad=True, fused=True)
        g = torch.Generator().manual_seed(0)
        x = torch.randn(Bsz, 4, S, S, S, generator=g).to(dev)
        t = (torch.rand(Bsz, 3, S, S, S, generator=g) > 0.7).float().to(dev)

        def step():
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                probs = O.unet_forward(leaves, [x])
            loss = O.dice_loss_joint([probs[0].float()], [t])
            loss.backward()
            opt.step()
        for _ in range(3):
            step()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        reps = 5
        for _ in range(reps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        del leaves, opt, x, t
        torch.cuda.empty_cache()
        return {"value": Bsz * S ** 3 / (ms * 1e-3), "unit": "voxels/s", "ms_per_step": ms, "kind": "stock PyTorch %s / cuDNN, "
                "bf16 autocast, fused Adam, eager; the reference arithmetic (oracle restatement of model.py + loss.py)" % torch.__version__}
    except Exception as e:
        return {"unavailable": repr(e)[:200]}


def ddp_parity(model, net, crit, rank, world, dev):
    """N ranks (one small volume each) against ONE process on the concatenated batch, same weights: the loss every rank
    reports and the 86 gradients after the bucketed all-reduce (tools/ddp_parity.py is the long form)."""
    import torch
    import torch.distributed as dist
    S = 32
    g = torch.Generator().manual_seed(4242)
    X = torch.randn(world, 4, S, S, S, generator=g).to(dev)
    T = (torch.rand(world, 3, S, S, S, generator=g) > 0.7).float().to(dev)
    model.zero_grad(set_to_none=True)
    loss = crit(net([X[rank:rank + 1].contiguous()]), [T[rank:rank + 1].contiguous()])
    loss.backward()
    torch.cuda.synchronize()
    losses = [torch.zeros(1, device=dev) for _ in range(world)]
    dist.all_gather(losses, loss.detach().reshape(1))
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    out = None
    if rank == 0:
        group, crit.process_group = crit.process_group, None
        factory, model._grad_store_factory = model._grad_store_factory, None
        model.zero_grad(set_to_none=True)
        l1 = crit(model([X]), [T])
        l1.backward()
        torch.cuda.synchronize()
        rel = [((grads[n] - p.grad).norm() / p.grad.norm().clamp_min(1e-20)).item() for n, p in model.named_parameters()
               if p.grad is not None]
        crit.process_group, model._grad_store_factory = group, factory
        ls = [float(v) for v in losses]
        out = {"shape": "%d ranks x 1 x 4x%d^3 vs 1 process on the concatenated batch" % (world, S),
               "loss_identical_across_ranks": max(ls) == min(ls), "loss_ddp": ls[0], "loss_single": float(l1),
               "grad_tensors": len(rel), "grad_rel_l2_max": max(rel), "grad_rel_l2_median": sorted(rel)[len(rel) // 2]}
    model.zero_grad(set_to_none=True)
    dist.barrier()
    return out


if __name__ == "__main__":
    main()
