"""Diagnostic: which part of the step breaks CUDA-graph capture."""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import brats2019_b200 as B  # noqa: E402
from brats2019_b200 import ops  # noqa: E402


def attempt(name, fn, warm):
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                warm()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out = fn()
        g.replay()
        torch.cuda.synchronize()
        print("CAPTURE OK  ", name, flush=True)
        return out
    except Exception as e:
        print("CAPTURE FAIL", name, "->", str(e).splitlines()[0][:160], flush=True)
        traceback.print_exc(limit=3)
        return None


def child(which):
    m = B.UNet(**B.DEFAULT_CFG).cuda().train()
    crit = B.Dice_loss_joint()
    opt = torch.optim.Adam(m.parameters(), lr=1e-4, amsgrad=True, capturable=True)
    x = torch.randn(1, 4, 16, 16, 16, device="cuda")
    t = (torch.rand(1, 3, 16, 16, 16, device="cuda") > 0.7).float()

    def fwd_nograd():
        with torch.no_grad():
            return m([x])[0]

    def fwd_grad():
        return m([x])[0]

    def fwd_loss():
        return crit(m([x]), [t])

    def manual_bwd():
        eng = m._engine()
        probs = eng.forward(x, training=True)
        sums = ops.dice_sums(probs, t)
        one = torch.ones(1, device="cuda")
        gp = ops.dice_backward(probs, t, sums, one, 1.0)
        return eng.backward(gp)

    def full_bwd():
        opt.zero_grad(set_to_none=True)
        loss = crit(m([x]), [t])
        loss.backward()
        return loss

    def full_step():
        opt.zero_grad(set_to_none=True)
        loss = crit(m([x]), [t])
        loss.backward()
        opt.step()
        return loss

    fn = dict(fwd_nograd=fwd_nograd, fwd_grad=fwd_grad, fwd_loss=fwd_loss, manual_bwd=manual_bwd, full_bwd=full_bwd,
              full_step=full_step)[which]
    attempt(which, fn, fn)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        child(sys.argv[1])
    else:
        import subprocess
        for w in ("fwd_nograd", "fwd_grad", "fwd_loss", "manual_bwd", "full_bwd", "full_step"):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), w], capture_output=True, text=True, timeout=300)
            lines = [ln for ln in (r.stdout + r.stderr).splitlines() if "CAPTURE" in ln or "Error" in ln or "File" in ln]
            print("\n".join(lines[:14]), flush=True)
