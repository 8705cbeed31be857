"""examples/c_abi_demo.c: the boundary used from plain C (gcc, no Python / torch in that process).  Without a GPU the
library must say so (APG_ERR_NO_DEVICE) instead of computing anything on the CPU; on a GPU the demo's loss is
checked against the oracle and its own finite-difference check of the analytic gradient must pass."""
import os
import re
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "apg_trajectory_tracking_b200")


@pytest.fixture(scope="module")
def demo(tmp_path_factory):
    if not os.path.exists(os.path.join(PKG, "libapg_b200.so")):
        from apg_trajectory_tracking_b200 import build as B
        B.build(force=False)
    exe = str(tmp_path_factory.mktemp("c_abi") / "c_abi_demo")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-o", exe, "-L", PKG, "-lapg_b200", "-lm",
                           "-Wl,-rpath," + PKG])
    return exe


def _lcg_inputs(n, h):
    """the demo's deterministic inputs (same LCG, same draw order)"""
    state = [12345]

    def u():
        state[0] = (state[0] * 1664525 + 1013904223) & 0xffffffff
        return np.float32(((state[0] >> 8) & 0xffffff) / 8388608.0 - 1.0)
    shapes = [(32, 4), (32,), (64, 32), (64,), (64, 64), (64,), (32, 64), (32,), (h, 32), (h,)]
    npar = sum(int(np.prod(s)) for s in shapes)
    params = np.zeros(npar, np.float32)
    for i in range(npar):
        params[i] = np.float32(0.2) * u()
        u()                                                     # the direction vector's draw
    st = np.zeros((n, 4), np.float32)
    for i in range(n):
        st[i] = [np.float32(2.4) * u(), np.float32(1.5) * u(), np.float32(0.5) * u(), np.float32(1.5) * u()]
    out, o = [], 0
    for s in shapes:
        k = int(np.prod(s))
        out.append(torch.tensor(params[o:o + k].reshape(s)))
        o += k
    return out, torch.tensor(st)


def test_c_demo_builds_and_refuses_to_run_without_a_gpu(demo):
    r = subprocess.run([demo], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 3 and "no CUDA device" in r.stderr and r.stdout == ""


@pytest.mark.gpu
def test_c_demo_loss_matches_oracle_and_gradient_check_passes(demo):
    from oracle import apg_oracle as O
    r = subprocess.run([demo], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    loss = float(re.search(r"^loss (\S+)", r.stdout, re.M).group(1))
    params, st = _lcg_inputs(256, 5)
    want, _, _, _ = O.concurrent_value_and_grad("cartpole", params, st, st, None, None, 5, 0.05)
    assert abs(loss - float(want)) <= 2e-5 * abs(float(want)), (loss, float(want))
