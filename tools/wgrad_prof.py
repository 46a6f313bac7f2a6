"""In-kernel cycle accounting of the weight-gradient GEMM (perf diagnostic, GPU box):
    python tools/wgrad_prof.py [B] [S] [C]
Runs the C x C 3x3x3 weight gradient with B200_WGRAD_DEBUG=256 and prints per-role averages over CTAs."""
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

os.environ["B200_WGRAD_DEBUG"] = "256"
import numpy as np  # noqa: E402
import torch  # noqa: E402

from brats2019_b200 import _lib, ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
Cc = int(sys.argv[3]) if len(sys.argv) > 3 else 16
dev = "cuda"
x = ops.act_zeros(B, S, S, S, Cc, dev)
dy = ops.act_zeros(B, S, S, S, Cc, dev)
for a in (x, dy):
    a.interior().copy_(torch.randn(Cc // 8, B, S, S, S, 8, device=dev).to(torch.bfloat16))
g = torch.empty(Cc, Cc, 3, 3, 3, device=dev)
wd = ops.wgrad_desc(0, B, S, S, S, Cc, Cc)
ws = ops.wgrad_workspace(wd, dev)
for _ in range(30):
    ops.wgrad_run(wd, dy, x, g, ops.G_K3, workspace=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.wgrad_run(wd, dy, x, g, ops.G_K3, workspace=ws)
e1.record()
torch.cuda.synchronize()
buf = (C.c_ulonglong * (160 * 16))()
_lib.check(_lib.lib().b200_march_prof_read(buf, 160 * 16), "prof")
a = np.array(list(buf), dtype=np.float64).reshape(160, 16)
a = a[a[:, 6] > 0]
names = ["prod_total", "prod_wait_empty", "prod_issue", "mma_total", "mma_wait_full", "mma_issue", "stages",
         "epi_wait_done", "epi_store"]
print("C=%d  %dx%d^3  ctas=%d  gemm+reduce %.1f us" % (Cc, B, S, len(a), e0.elapsed_time(e1) * 1e3))
for i, n in enumerate(names):
    print("  %-16s mean %10.0f  max %10.0f" % (n, a[:, i].mean(), a[:, i].max()))
