"""The rational tanh of the tcgen05 forward epilogue (csrc/tq_kernels.cu tq_tanh_rational), restated in numpy float32
with the SAME coefficients (read from the source file), against fp64 tanh: relative error over the whole range."""
import os
import re

import numpy as np

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "apg_trajectory_tracking_b200", "csrc",
                   "tq_kernels.cu")


def _coefficients():
    body = open(SRC).read()
    body = body[body.index("float tq_tanh_rational(float x)"):]
    body = body[:body.index("return (x * p)")]
    nums = [float(m) for m in re.findall(r"(-?\d\.\d+e-\d+)f", body)]
    clamp = float(re.search(r"fmaxf\(x, (-\d\.\d+)f\)", body).group(1))
    assert len(nums) == 11
    return -clamp, nums[:7], nums[7:]


def test_rational_tanh_relative_error_below_5e_7_everywhere():
    f = np.float32
    clamp, a, b = _coefficients()                      # a: x^12 .. x^0 of P, b: x^6 .. x^0 of Q
    x = np.concatenate([np.linspace(-12, 12, 1000001), np.logspace(-8, 0, 100001), -np.logspace(-8, 0, 100001)]).astype(f)
    xc = np.clip(x, f(-clamp), f(clamp))
    x2 = xc * xc
    p = np.full_like(x2, f(a[0]))
    for c in a[1:]:
        p = (p * x2 + f(c)).astype(f)
    q = np.full_like(x2, f(b[0]))
    for c in b[1:]:
        q = (q * x2 + f(c)).astype(f)
    y = ((xc * p).astype(f) / q).astype(f)
    t = np.tanh(x.astype(np.float64))
    rel = np.abs(y - t) / np.maximum(np.abs(t), 1e-300)
    assert rel.max() <= 5e-7, rel.max()
    assert np.abs(y).max() <= 1.0
