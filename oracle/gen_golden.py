"""Generate golden vectors from the REAL reference (run in the build container only).

    python oracle/gen_golden.py            # writes tests/golden/*.npz

Imports /root/reference/{model,loss}.py unmodified, loads parameters produced by
`oracle.resunet_oracle.init_params(seed)` (deterministic torch CPU generator) through the
reference's own `load_state_dict`, runs the reference forward / loss / autograd backward on
small seeded inputs and stores the results.  /root/reference does not exist on the GPU box,
so the vectors are committed; this script is the record of how they were made.
"""
import hashlib
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, "/root/reference")

from oracle import resunet_oracle as O  # noqa: E402

CASES = {
    # name: (batch, D, H, W, weight seed, data seed, target seed)
    "cube16_b1": (1, 16, 16, 16, 1337, 0, 1),
    "box16x24x32_b2": (2, 16, 24, 32, 1337, 10, 11),
}
FULL_GRADS = [
    "conv_input.weight", "norm_input.weight", "norm_input.bias", "conv_output.weight",
    "conv_output.bias", "conv_first.0.conv1.conv1.weight", "conv_first.0.norm2.weight",
    "encoder_convs.0.0.downsample.0.weight", "upsampling.0.1.weight",
    "decoder_convs1x1.0.weight", "decoder_convs.0.0.norm1.bias", "encoder_convs.2.3.norm2.weight",
    "upsampling.2.1.weight", "decoder_convs1x1.2.weight",
]


def make_inputs(b, d, h, w, data_seed, target_seed):
    g = torch.Generator().manual_seed(data_seed)
    x = torch.randn(b, 4, d, h, w, generator=g)
    g = torch.Generator().manual_seed(target_seed)
    t = (torch.rand(b, 3, d, h, w, generator=g) > 0.7).float()
    return x, t


def params_digest(sd):
    m = hashlib.sha256()
    for k, v in sd.items():
        m.update(k.encode())
        m.update(v.numpy().tobytes())
    return m.hexdigest()


def main():
    import model as ref_model   # /root/reference/model.py
    import loss as ref_loss     # /root/reference/loss.py

    out_dir = os.path.join(REPO, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (b, d, h, w, wseed, dseed, tseed) in CASES.items():
        sd = O.init_params(wseed)
        net = ref_model.UNet(**O.DEFAULT_CFG)
        assert [k for k, _ in O.param_shapes()] == list(net.state_dict().keys()), "name/order contract"
        net.load_state_dict(sd)
        net.train()
        x, t = make_inputs(b, d, h, w, dseed, tseed)

        captured = {}
        hook = net.conv_output.register_forward_hook(lambda m, i, o: captured.__setitem__("logits", o.detach()))
        probs = net([x])
        hook.remove()
        dice = ref_loss.Dice_loss_joint(index=0, priority=1)(probs, [t])
        bce = ref_loss.BCE_Loss(index=0, bg_weight=1e-2)(probs, [t])
        net.zero_grad()
        dice.backward(retain_graph=True)
        grads_dice = {k: (p.grad.clone() if p.grad is not None else None) for k, p in net.named_parameters()}
        net.zero_grad()
        ((dice + bce) / 2).backward()                      # train.py:203-210 with main.py:126-128
        grads_both = {k: (p.grad.clone() if p.grad is not None else None) for k, p in net.named_parameters()}

        dead = [k for k, g in grads_dice.items() if g is None]
        assert sorted(dead) == sorted(O.dead_param_names()), dead
        live = [k for k in grads_dice if grads_dice[k] is not None]
        rec = dict(
            x=x.numpy(), target=t.numpy().astype(np.uint8),
            logits=captured["logits"].numpy(), probs=probs[0].detach().numpy(),
            dice=np.float64(dice.item()), bce=np.float64(bce.item()),
            params_sha256=np.array(params_digest(sd)),
            weight_seed=np.int64(wseed),
            live_names=np.array(live),
            grad_norm_dice=np.array([grads_dice[k].double().norm().item() for k in live]),
            grad_sum_dice=np.array([grads_dice[k].double().sum().item() for k in live]),
            grad_norm_both=np.array([grads_both[k].double().norm().item() for k in live]),
        )
        for k in FULL_GRADS:
            rec["gdice::" + k] = grads_dice[k].numpy()
            rec["gboth::" + k] = grads_both[k].numpy()
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **rec)
        print(name, "dice", dice.item(), "bce", bce.item(), "->", path, os.path.getsize(path) // 1024, "KiB")


def tiling_golden():
    """Pin the oracle's restatement of the sliding-window arithmetic on the reference's own
    loader_helper.get_indices/copy/copy_back (nibabel is absent here: stubbed, it is only used by
    the NIfTI readers of that file)."""
    import types
    sys.modules.setdefault("nibabel", types.ModuleType("nibabel"))
    import loader_helper as LH           # /root/reference/loader_helper.py
    g = torch.Generator().manual_seed(3)
    data = torch.randn(1, 2, 21, 13, 10, generator=g)
    tile, center, border = (10, 8, 6), (4, 4, 2), (3, 2, 2)
    out = torch.zeros(1, 2, 21, 13, 10)
    grid = [int(np.ceil(j / i)) for i, j in zip(center, data.shape[2:])]
    tiles = []
    for i in range(grid[0]):
        for j in range(grid[1]):
            for k in range(grid[2]):
                imin, imax = LH.get_indices(position=(i, j, k), center_shape=center, border=border)
                t = LH.copy(data=data, tile_shape=tile, index_min=imin, index_max=imax)
                tiles.append(t.numpy().copy())
                f = t * 2.0 + (i * 100 + j * 10 + k)          # stand-in for the model: position-dependent
                LH.copy_back(data=out, tile=f, center_shape=center, index_min=imin, index_max=imax, border=border)
    path = os.path.join(REPO, "tests", "golden", "tiling.npz")
    np.savez_compressed(path, data=data.numpy(), out=out.numpy(), tiles=np.stack(tiles), tile=np.array(tile),
                        center=np.array(center), border=np.array(border))
    print("tiling ->", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    torch.manual_seed(0)
    if len(sys.argv) > 1 and sys.argv[1] == "tiling":
        tiling_golden()
    else:
        main()
        tiling_golden()
