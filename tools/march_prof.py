"""In-kernel cycle accounting of the marching conv (perf diagnostic, GPU box):
    python tools/march_prof.py [B] [S] [C] [debugflags]
Runs the C->C 3x3x3 conv with B200_CONV_DEBUG = 256 | flags and prints per-role averages over CTAs."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
# the skip-a-stage / cycle-accounting code exists only in the probes build: python -m brats2019_b200.build --probes
_PROBES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "brats2019_b200", "libbrats_b200_probes.so")
if True:
    if not os.path.exists(_PROBES):
        raise SystemExit("build the probes library first: python -m brats2019_b200.build --probes")
    os.environ["B200_LIB_PATH"] = _PROBES

flags = int(sys.argv[4]) if len(sys.argv) > 4 else 0
os.environ["B200_CONV_DEBUG"] = str(256 | flags)
import torch  # noqa: E402

from brats2019_b200 import _lib, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
Cc = int(sys.argv[3]) if len(sys.argv) > 3 else 16
dev = "cuda"
x = ops.act_zeros(B, S, S, S, Cc, dev)
x.interior().copy_(torch.randn(Cc // 8, B, S, S, S, 8, device=dev).to(torch.bfloat16))
y = ops.act_zeros(B, S, S, S, Cc, dev)
w = torch.randn(Cc, Cc, 3, 3, 3, device=dev) * 0.05
desc = ops.conv_desc(ops.MODE_K3, B, S, S, S, Cc, Cc)
pk = ops.conv_pack_weight(desc, ops.W_FWD, w)
ctas = ops.conv_ctas(desc)
st = torch.empty(ctas * B * 16, device=dev)
for _ in range(30):      # also ramps the SM clock before the measured launch
    ops.conv_run(desc, x, pk, y, stats=st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.conv_run(desc, x, pk, y, stats=st)
e1.record()
torch.cuda.synchronize()
buf = (C.c_ulonglong * (160 * 16))()
_lib.check(_lib.lib().b200_march_prof_read(buf, 160 * 16), "prof")
import numpy as np  # noqa: E402
a = np.array(list(buf), dtype=np.float64).reshape(160, 16)[:ctas]
names = ["prod_total", "prod_wait_empty", "mma_total", "mma_wait_x", "mma_wait_acc", "mma_issue", "steps",
         "g0_wait_done", "g0_ld", "g0_st", "g0_rest", "g1_wait_done", "g1_ld", "g1_st", "g1_rest", "epi_total"]
print("flags=%d  C=%d  %dx%d^3  ctas=%d  kernel %.1f us" % (flags, Cc, B, S, ctas, e0.elapsed_time(e1) * 1e3))
for i, n in enumerate(names):
    print("  %-16s mean %10.0f  max %10.0f" % (n, a[:, i].mean(), a[:, i].max()))
if a[:, 0].mean() > 0:
    print("  (band kernel) mma warp span %.1f us -> SM clock %.0f MHz; CTA start skew %.1f us; end spread %.1f us" % (
        a[:, 0].mean() / 1e3, a[:, 2].mean() / a[:, 0].mean() * 1e3, (a[:, 1].max() - a[:, 1].min()) / 1e3,
        ((a[:, 1] + a[:, 0]).max() - (a[:, 1] + a[:, 0]).min()) / 1e3))
