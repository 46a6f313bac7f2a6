"""Compact host staging formats (bf16 images, uint8 / bool targets) give bit-identical results to the reference's fp32
tensors: the packing kernel rounds images to bf16 anyway, and 0/1 targets are exact in every format."""
import pytest
import torch

from oracle import resunet_oracle as O

pytestmark = pytest.mark.gpu


def test_bf16_images_and_uint8_targets_are_bit_identical_to_fp32():
    import brats2019_b200 as B
    m = B.UNet(**B.DEFAULT_CFG)
    m.load_state_dict(O.init_params(1337))
    m = m.cuda().train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 4, 16, 24, 40, generator=g).cuda()
    t = (torch.rand(2, 3, 16, 24, 40, generator=g) > 0.7).float().cuda()
    crits = [B.Dice_loss_joint(), B.BCE_Loss(bg_weight=1e-2)]

    def run(xx, tt):
        m.zero_grad(set_to_none=True)
        out = m([xx])
        losses = [c(out, [tt]) for c in crits]
        (sum(losses) / 2).backward()
        return out[0].detach().clone(), [l.item() for l in losses], {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}

    p0, l0, g0 = run(x, t)
    for xx, tt in ((x.bfloat16(), t.to(torch.uint8)), (x.bfloat16(), t.bool()), (x, t.to(torch.uint8))):
        p1, l1, g1 = run(xx, tt)
        assert torch.equal(p0, p1) and l0 == l1
        assert all(torch.equal(g0[k], g1[k]) for k in g0)
    # odd W (scalar packing path) and a misaligned bf16 view
    xo = torch.randn(1, 4, 8, 8, 24, generator=g).cuda()
    xb = torch.empty(xo.numel() + 1, dtype=torch.bfloat16, device="cuda")[1:].view_as(xo).copy_(xo)
    m.eval()
    with torch.no_grad():
        assert torch.equal(m([xo])[0], m([xb])[0])
