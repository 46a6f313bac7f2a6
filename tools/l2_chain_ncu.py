"""One eager pass of the conv -> finalize -> apply chain for ncu (--cache-control none): python tools/l2_chain_ncu.py N S"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from brats2019_b200 import ops  # noqa: E402

N = int(sys.argv[1]); S = int(sys.argv[2]); K = 3; Cc = 16; dev = "cuda"
mk = lambda: ops.act_zeros(N, S, S, S, Cc, dev)
acts = [mk() for _ in range(K + 1)]
convs = [mk() for _ in range(K)]
acts[0].interior().copy_(torch.randn(Cc // 8, N, S, S, S, 8, device=dev).to(torch.bfloat16))
desc = ops.conv_desc(ops.MODE_K3, N, S, S, S, Cc, Cc)
w = torch.randn(Cc, Cc, 3, 3, 3, device=dev) / (Cc * 27) ** 0.5
pk = ops.conv_pack_weight(desc, ops.W_FWD, w, K_real=Cc, N_real=Cc)
ctas = ops.conv_ctas(desc)
stats = torch.empty(ctas * N * 16, device=dev)
mean = torch.empty(N * 8, device=dev); rstd = torch.empty(N * 8, device=dev)
gamma = torch.ones(Cc, device=dev); beta = torch.zeros(Cc, device=dev)
for rep in range(2):
    for k in range(K):
        ops.conv_run(desc, acts[k], pk, convs[k], stats=stats)
        ops.gn_finalize(stats, ctas, N, Cc, S, S, S, mean, rstd)
        ops.gn_apply(convs[k], mean, rstd, gamma, beta, acts[k + 1], lrelu=True)
torch.cuda.synchronize()
