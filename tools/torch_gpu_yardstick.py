"""Speed yardstick (diagnostic, run on the GPU box): the REFERENCE ARITHMETIC as stock PyTorch on the same
B200 - the oracle restatement on CUDA, i.e. cuDNN convolutions + ATen GroupNorm/upsample - in fp32, TF32 and
bf16 autocast, at the bench workload (train step on B x 4xS^3, forward on the same batch).  SURVEY.md 8d asks
for this next to the CPU baseline; it is what a user of the reference gets by just moving to a B200.

    python tools/torch_gpu_yardstick.py [B] [S] [reps]
Prints one JSON line per mode: ms per train step (fwd + Dice + autograd backward, no optimizer) and per forward."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402

from oracle import resunet_oracle as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
sd = {k: v.cuda() for k, v in O.init_params(1337).items()}
g = torch.Generator().manual_seed(0)
x = torch.randn(B, 4, S, S, S, generator=g).cuda()
t = (torch.rand(B, 3, S, S, S, generator=g) > 0.7).float().cuda()
torch.backends.cudnn.benchmark = True          # main.py:73


def timeit(fn):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def fwd():
    with torch.no_grad():
        O.unet_forward(sd, [x])


for mode in ("fp32", "tf32", "bf16-autocast"):
    torch.backends.cudnn.allow_tf32 = mode == "tf32"
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode == "bf16-autocast" else torch.autocast("cuda", enabled=False)
    with ctx:
        ms_step = timeit(lambda: O.train_step(sd, x, t))
        ms_fwd = timeit(fwd)
    vox = B * S ** 3
    print(json.dumps({"impl": "stock PyTorch %s on this GPU (cuDNN %s), reference arithmetic" % (
        torch.__version__, torch.backends.cudnn.version()), "mode": mode, "batch": B, "size": S,
        "train_step_ms": ms_step, "train_voxels_per_s": vox / (ms_step * 1e-3),
        "forward_ms": ms_fwd, "forward_voxels_per_s": vox / (ms_fwd * 1e-3),
        "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9}), flush=True)
