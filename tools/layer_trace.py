"""Layer-by-layer comparison (diagnostic, GPU box): every stored activation of the CUDA engine against the
bf16-storage emulation of the oracle, in execution order.  python tools/layer_trace.py [N D H W]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import torch  # noqa: E402

from oracle import resunet_oracle as O  # noqa: E402


def main():
    import brats2019_b200 as B
    from brats2019_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    a = [int(v) for v in sys.argv[1:]] or [1, 32, 32, 32]
    N, D, H, W = a
    sd = O.init_params(1337)
    sdc = {k: v.cuda() for k, v in sd.items()}
    g = torch.Generator().manual_seed(7)
    x = torch.randn(N, 4, D, H, W, generator=g).cuda()
    trace = {}
    with torch.no_grad():
        logits_em = O.unet_logits_bf16_emulated(sdc, x, trace=trace)
        trace32 = {}
        with O.rounding(False):
            O.unet_logits_bf16_emulated(sdc, x, trace=trace32)
    m = B.UNet(**B.DEFAULT_CFG)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    (_,), logits = m([x], return_logits=True)
    P = list(m._engine().plans.values())[0]
    acts = {k[0]: v for k, v in P.acts.items()}
    print("%-36s %10s %10s %10s %10s" % ("tensor", "rel-L2 em", "max/max em", "neq frac", "rel-L2 f32"))
    for name, ref in trace.items():
        if name not in acts:
            print("%-36s (no engine buffer)" % name)
            continue
        got = ops.act_to_ncdhw(acts[name])
        r32 = trace32[name]
        print("%-36s %10.2e %10.2e %10.5f %10.2e" % (name, ((got - ref).norm() / ref.norm()).item(),
                                                   ((got - ref).abs().max() / ref.abs().max()).item(),
                                                   (got != ref).float().mean().item(), ((got - r32).norm() / r32.norm()).item()))
    print("logits rel-L2 vs emulated %.3e" % ((logits - logits_em).norm() / logits_em.norm()).item())


if __name__ == "__main__":
    main()
