"""Shared helpers for the test-suite (golden loading, tolerances)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return {k: z[k] for k in z.files}


def golden_params(g, dtype=torch.float32):
    n = len(g["param_names"])
    return [torch.tensor(g[f"param_{i}"], dtype=dtype) for i in range(n)]


def golden_grads(g):
    n = len(g["param_names"])
    return [torch.tensor(g[f"grad_{i}"]) if f"grad_{i}" in g else None for i in range(n)]


def t(x, dtype=torch.float32):
    return torch.tensor(np.asarray(x), dtype=dtype)


def _cpu64(x):
    if isinstance(x, torch.Tensor):
        return x.detach().to(device="cpu", dtype=torch.float64)
    return torch.as_tensor(x, dtype=torch.float64)


def rel_err(a, b):
    a = _cpu64(a)
    b = _cpu64(b)
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_rel_to_scale(a, b):
    """max |a-b| / max|b|  -- elementwise error relative to the tensor's scale."""
    a = _cpu64(a)
    b = _cpu64(b)
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))
