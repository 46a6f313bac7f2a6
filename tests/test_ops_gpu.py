"""GPU: every kernel family behind the C ABI against a plain torch fp32 reference of the same
op (the checks live in tests/gpu_opcheck.py so they can also run as a diagnostic script)."""
import importlib.util
import os

import pytest

pytestmark = pytest.mark.gpu

_spec = importlib.util.spec_from_file_location("gpu_opcheck", os.path.join(os.path.dirname(__file__), "gpu_opcheck.py"))
opcheck = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(opcheck)


@pytest.mark.parametrize("group", ["ew", "conv3", "conv1", "dgrad", "wgrad"])
def test_op_group(group):
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    del opcheck.RESULTS[:]
    opcheck.GROUPS[group]()
    torch.cuda.synchronize()
    assert opcheck.RESULTS and all(opcheck.RESULTS), "%d of %d checks failed in group %s" % (
        sum(1 for r in opcheck.RESULTS if not r), len(opcheck.RESULTS), group)
