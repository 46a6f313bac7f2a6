#!/usr/bin/env python
"""Inference throughput of the drop-in UNet at the BraTS shapes (BASELINE.json configs 2 and 5).

    python tools/bench_inference.py [--volumes V] [--tiled-volumes T] [--reps R]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_inference.py ...

Config 2: one 4x240x240x155 volume zero-padded to 4x240x240x160 (test.py:93), eval / no_grad, B = 1.
Config 5: V volumes per GPU, each 4-flip TTA (test.py:115-141) -> threshold -> label painting, (a) whole
volume, which is what test.py does, and (b) through the sliding window of Trainer.predict_tiled
(train.py:145-176: tile 192^3, centre 48^3 -> 5x5x4 = 100 tiles per pass).  Volumes are sharded across
ranks (parallel.shard_volumes), no collective on the data path; times are CUDA events, max over ranks.
Prints one JSON line per measurement on rank 0.  Synthetic data, random-init weights.
"""
import argparse
import contextlib
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

FLOP_PER_VOXEL_FWD = 142752


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--volumes", type=int, default=4, help="whole-volume TTA volumes per GPU")
    ap.add_argument("--tiled-volumes", type=int, default=1, help="sliding-window TTA volumes per GPU")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--batch-tiles", type=int, default=2)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import brats2019_b200 as B
    from brats2019_b200 import inference as I
    from brats2019_b200.parallel import shard_volumes

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(REPO, "MEASURED_PEAKS.json")) else {"bf16_tflops_sustained": 1400.0}

    def sync_max(ms):
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        return sync_max(e0.elapsed_time(e1)) / reps

    def emit(d):
        if rank == 0:
            d.update({"n_gpus": world, "dtype": "bf16", "data": "synthetic"})
            print(json.dumps(d), flush=True)

    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream), torch.no_grad():
        torch.manual_seed(1337)
        with contextlib.redirect_stdout(sys.stderr):
            model = B.UNet(**B.DEFAULT_CFG).to(dev).eval()
        g = torch.Generator().manual_seed(rank)
        raw = torch.randn(1, 4, 240, 240, 155, generator=g).to(dev)
        x, left, right = I.pad_to_multiple(raw, 16)                      # -> 240x240x160
        vox = x.shape[2] * x.shape[3] * x.shape[4]

        # ---- config 2: single forward, B = 1 ----
        for _ in range(3):
            model([x])
        ms = timed(lambda: model([x]), args.reps)
        emit({"config": "2: whole-volume forward 1 x 4x240x240x160 per GPU", "ms_per_volume": ms,
              "voxels_per_s": vox * world / (ms * 1e-3),
              "tensor_frac_of_sustained": vox / (ms * 1e-3) * FLOP_PER_VOXEL_FWD / 1e12 / peaks["bf16_tflops_sustained"]})

        # ---- config 5a: whole-volume 4-flip TTA + threshold + label painting, V volumes per GPU ----
        def whole(v):
            p = I.predict_tta(model, x)
            return I.paint_labels(I.unpad(p, left, right))

        whole(0)
        nv = args.volumes
        mine = len(shard_volumes(nv * world, rank, world))
        ms = timed(lambda: [whole(v) for v in range(mine)], 1)
        emit({"config": "5a: %d volumes, 4-flip TTA whole volume (test.py:115-159), sharded by volume" % (nv * world),
              "ms_total": ms, "volumes_per_s": nv * world / (ms * 1e-3),
              "output_voxels_per_s": nv * world * 240 * 240 * 155 / (ms * 1e-3),
              "model_voxels_per_s": 4 * nv * world * vox / (ms * 1e-3)})

        # ---- config 5b: the same through the 192^3 sliding window ----
        if args.tiled_volumes > 0:
            def tiled_tta():
                acc = None
                for dims in ((), (2,), (3,), (2, 3)):
                    xi = torch.flip(x, dims).contiguous() if dims else x
                    p = I.predict_tiled(model, xi, batch_tiles=args.batch_tiles)
                    p = torch.flip(p, dims) if dims else p
                    acc = p if acc is None else acc.add_(p)
                return I.paint_labels(I.unpad(acc / 4.0, left, right))

            I.predict_tiled(model, x, batch_tiles=args.batch_tiles)      # warm-up: plans for the tile shape
            tv = args.tiled_volumes
            ms = timed(lambda: [tiled_tta() for _ in range(tv)], 1)
            tiles = 4 * 100 * tv * world
            emit({"config": "5b: %d volumes, 4-flip TTA x sliding window (tile 192^3, centre 48^3, 100 tiles/pass, "
                            "%d tiles per batch)" % (tv * world, args.batch_tiles),
                  "ms_total": ms, "volumes_per_s": tv * world / (ms * 1e-3),
                  "output_voxels_per_s": tv * world * 240 * 240 * 155 / (ms * 1e-3),
                  "tile_voxels_per_s": tiles * 192 ** 3 / (ms * 1e-3),
                  "tensor_frac_of_sustained": tiles / world * 192 ** 3 / (ms * 1e-3) * FLOP_PER_VOXEL_FWD / 1e12
                  / peaks["bf16_tflops_sustained"]})
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
