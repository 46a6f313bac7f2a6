"""Parity report (diagnostic, run on the GPU box):  CUDA ResUNet vs the fp32 oracle, next to the
yardstick "the reference arithmetic itself in bf16" (the oracle under torch.autocast(bfloat16)
on the same GPU, i.e. stock PyTorch/cuDNN bf16), on the same weights and inputs.

    python tools/parity_report.py [N D H W]...     # default: a few small/medium shapes

Third line per shape: the CUDA path against the bf16-storage EMULATION of itself
(oracle.train_step_bf16_emulated: the reference algorithm with roundings where the kernels store) - the
comparison that is not limited by the bf16 noise floor.  `profiles/r02_parity_report.txt` is a committed copy.
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

import torch  # noqa: E402

from oracle import resunet_oracle as O  # noqa: E402


def stats_forward(logits, ref):
    err = (logits - ref).abs()
    p, pr = torch.sigmoid(logits), torch.sigmoid(ref)
    ma, mb = p > 0.5, pr > 0.5
    dice = []
    for c in range(p.shape[1]):
        den = ma[:, c].sum().item() + mb[:, c].sum().item()
        dice.append(2.0 * (ma[:, c] & mb[:, c]).sum().item() / den if den else 1.0)
    return dict(max_rel=err.max().item() / ref.abs().max().item(), mean_rel=err.mean().item() / ref.abs().mean().item(),
                prob_max=(p - pr).abs().max().item(), flipped=(ma != mb).float().mean().item(), mask_dice=min(dice))


def stats_grads(grads, ref):
    rels = []
    for k, r in ref.items():
        g = grads[k].float().to(r.device)
        rels.append(((g - r).norm() / r.norm().clamp_min(1e-20)).item())
    t = torch.tensor(rels)
    return dict(max=t.max().item(), median=t.median().item(), worst=list(ref.keys())[int(t.argmax())])


def main():
    import brats2019_b200 as B
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    args = [int(a) for a in sys.argv[1:]]
    shapes = [tuple(args[i:i + 4]) for i in range(0, len(args), 4)] or [(1, 16, 16, 16), (2, 16, 24, 32), (1, 32, 32, 32),
                                                                       (2, 32, 48, 64), (1, 64, 64, 64)]
    sd = O.init_params(1337)
    sd_cuda = {k: v.cuda() for k, v in sd.items()}
    for (N, D, H, W) in shapes:
        g = torch.Generator().manual_seed(7)
        x = torch.randn(N, 4, D, H, W, generator=g).cuda()
        t = (torch.rand(N, 3, D, H, W, generator=g) > 0.7).float().cuda()
        # fp32 reference arithmetic on the GPU (cuDNN fp32, TF32 off)
        loss_ref, _, grads_ref = O.train_step(sd_cuda, x, t)
        logits_ref = O.unet_logits(sd_cuda, x)
        # yardstick: same arithmetic under bf16 autocast
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss_bf, _, grads_bf = O.train_step(sd_cuda, x, t)
            logits_bf = O.unet_logits(sd_cuda, x).float()
        # ours
        m = B.UNet(**B.DEFAULT_CFG)
        m.load_state_dict(sd)
        m = m.cuda()
        (_,), logits = m([x], return_logits=True)
        out = m([x])
        loss = B.Dice_loss_joint()(out, [t])
        loss.backward()
        grads = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
        fo, fb = stats_forward(logits, logits_ref), stats_forward(logits_bf, logits_ref)
        go, gb = stats_grads(grads, grads_ref), stats_grads(grads_bf, grads_ref)
        loss_em, probs_em, logits_em, grads_em = O.train_step_bf16_emulated(sd_cuda, x, t)
        fe, ge = stats_forward(logits, logits_em), stats_grads(grads, grads_em)
        dm = (O.dice_metric(out[0].detach() > 0.5, t > 0.5) - O.dice_metric(torch.sigmoid(logits_ref) > 0.5, t > 0.5)).abs().max().item()
        print("shape N=%d %dx%dx%d  |logit| max %.2f mean %.3f  dice loss ref %.6f" % (
            N, D, H, W, logits_ref.abs().max().item(), logits_ref.abs().mean().item(), loss_ref.item()))
        for tag, f, gg, l in (("b200 kernels ", fo, go, loss.item()), ("torch bf16 AMP", fb, gb, loss_bf.item())):
            print("  %s logits: max-err/max %.4f  mean-err/mean %.4f | prob max-err %.4f | mask flipped %.5f dice %.5f | "
                  "loss diff %.2e | grad rel-L2 max %.4f median %.4f (%s)" % (
                      tag, f["max_rel"], f["mean_rel"], f["prob_max"], f["flipped"], f["mask_dice"],
                      abs(l - loss_ref.item()), gg["max"], gg["median"], gg["worst"]))
        print("  b200 vs bf16-emulated oracle: logits max-err/max %.5f mean-err/mean %.5f | prob max-err %.5f | mask flipped %.6f "
              "dice %.6f | loss diff %.2e | grad rel-L2 max %.5f median %.5f (%s) | metrics.Dice(mask,target) diff vs fp32 ref %.2e" % (
                  fe["max_rel"], fe["mean_rel"], fe["prob_max"], fe["flipped"], fe["mask_dice"], abs(loss.item() - loss_em.item()),
                  ge["max"], ge["median"], ge["worst"], dm))
        sys.stdout.flush()
        del m


if __name__ == "__main__":
    main()
