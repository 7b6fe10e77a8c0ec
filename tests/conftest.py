import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
SHADE_CASES = ["small", "k50", "uneven", "empty"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    """tests/golden/<name>.npz -> dict of torch tensors / python scalars (made by make_golden.py
    from the unmodified reference)."""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {}
    for k in z.files:
        v = z[k]
        out[k] = torch.from_numpy(v) if v.ndim > 0 else v.item()
    return out


@pytest.fixture(params=SHADE_CASES)
def shade_case(request):
    g = load_golden("shade_" + request.param)
    g["name"] = request.param
    g["a_s"] = g["a_s"].long()
    g["a_0"] = g["a_0"].long()
    g["znear_t"] = g["znear"].reshape(-1, 1, 1, 1)
    g["zfar_t"] = g["zfar"].reshape(-1, 1, 1, 1)
    return g


def rel_err(a, b):
    """max|a-b| / max(|b|, tiny): norm-relative error used for gradient parity."""
    a, b = a.double(), b.double()
    denom = b.abs().max().clamp(min=1e-30)
    return ((a - b).abs().max() / denom).item()


def elementwise_close(a, b, rtol, floor=0.05):
    """Element-wise relative agreement |a-b| <= rtol * max(|b|, floor * max|b|): north_star's "1e-5 relative error on images
    and gradients" entry by entry, with an absolute floor (a fraction of the largest magnitude) so that entries which
    cancel to nearly zero are not compared at a precision their own terms do not have."""
    a, b = a.double(), b.double()
    scale = b.abs().max().clamp(min=1e-30)
    return bool(((a - b).abs() <= rtol * torch.maximum(b.abs(), floor * scale)).all())
