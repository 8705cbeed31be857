"""Control flow of FusedTrainStep.step_host (chunking, staging, loss / gradient accumulation over the chunks, the
per-system input derivation) exercised WITHOUT a GPU: the CUDA stream / event objects are replaced by no-ops, the
rollout runner by a stand-in that evaluates the CPU oracle, and the device prepare ops by the host dataset mirror.
Nothing of the product computes on the CPU -- this is a test double for host-side logic only; the real kernels
behind the same code are checked by tests/test_zz_new_paths_gpu.py on the GPU."""
import contextlib
import ctypes

import pytest
import torch

import bench as B
from apg_trajectory_tracking_b200 import _capi, prepare as PR, rollout as R, synthetic as SY, train as T
from apg_trajectory_tracking_b200.neural_control import dataset as DS
from oracle import apg_oracle as O
from tests.helpers import rel_err


class _Stream:
    def __init__(self, *a, **k):
        pass

    def wait_event(self, ev):
        assert ev.recorded, "waiting on an event that was never recorded"

    def wait_stream(self, other):
        pass


class _Event:
    def __init__(self):
        self.recorded = False

    def record(self, stream=None):
        self.recorded = True

    def synchronize(self):
        assert self.recorded


class _OracleRunner:
    """stands in for rollout.Rollout: same attributes / calls, oracle math on CPU tensors"""
    made = []

    def __init__(self, spec, n, device=None):
        self.spec, self.n, self.device = spec, int(n), torch.device("cpu")
        cfg = spec.config(self.n)
        self.n_params = _capi.lib().apg_num_params(ctypes.byref(cfg))      # host-only layout arithmetic
        self.loss = torch.zeros(1)
        _OracleRunner.made.append(self.n)

    def forward(self, flat, in_state, cur, in_ref=None, ref=None, h0c0=None, **kw):
        s = self.spec
        assert cur.shape[0] == self.n, "chunk size does not match the runner"
        shapes = B.hutter_shapes(s.system, s.horizon, s.mode if s.mode != "LSTM" else "lstm")
        like = [torch.empty(*sh) for sh in shapes]
        params = [p.clone() for p in R.split_flat(flat, like)]
        if s.mode == "concurrent":
            loss, grads, _, _ = O.concurrent_value_and_grad(s.system, params, in_state, cur, in_ref, ref, s.horizon,
                                                            s.dt)
        else:
            hc = (h0c0[0], h0c0[1]) if h0c0 is not None else None
            loss, grads, _, _ = O.recurrent_value_and_grad(s.mode.lower(), params, cur, in_ref, ref, s.horizon, s.dt,
                                                           hc0=hc)
        self._g = torch.cat([(g if g is not None else torch.zeros_like(p)).reshape(-1) for g, p in zip(grads, params)])
        self.loss[0] = float(loss)
        return self.loss, None, None

    def backward(self, grad_loss=1.0, out=None):
        if out is None:
            out = torch.empty_like(self._g)
        out.copy_(self._g * grad_loss)
        return out

    def value_and_grad(self, flat, in_state, cur, in_ref=None, ref=None, h0c0=None, out=None):
        loss, _, _ = self.forward(flat, in_state, cur, in_ref, ref, h0c0)
        return loss, self.backward(1.0, out=out)


def _prepare_quad(states, ref_states, want=(), out=None, in_place=False):
    ins, cur, inr, ref = DS.QuadDataset.prepare_data(None, states.clone(), ref_states.clone())
    res = {"in_state": ins, "cur": cur, "in_ref": inr, "ref": ref}
    for k in want:
        out[k].copy_(res[k])
    return out


def _prepare_wing(states, targets, mean, std, dt, horizon, want=("in_state", "cur", "in_ref", "ref"), out=None):
    ds = DS.WingDataset.__new__(DS.WingDataset)
    ds.dt, ds.horizon, ds.mean, ds.std = dt, horizon, torch.as_tensor(mean).float(), torch.as_tensor(std).float()
    ins, cur, inr, ref = ds.prepare_data(states.clone(), targets.clone())
    res = {"in_state": ins, "cur": cur, "in_ref": inr, "ref": ref}
    for k in want:
        if out[k].data_ptr() != states.data_ptr():
            out[k].copy_(res[k])
    return out


@pytest.fixture
def cpu_doubles(monkeypatch):
    monkeypatch.setattr(R, "Rollout", _OracleRunner)
    monkeypatch.setattr(PR, "prepare_quad", _prepare_quad)
    monkeypatch.setattr(PR, "prepare_wing", _prepare_wing)
    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(T.FusedTrainStep, "default_chunk", lambda self: 128)
    _OracleRunner.made = []


CASES = {
    "quad": dict(h=10, dt=0.1, n=200, chunk=64),
    "wing": dict(h=20, dt=0.05, n=150, chunk=64),
    "cartpole": dict(h=5, dt=0.05, n=130, chunk=128),
    "autoregressive": dict(h=4, dt=0.1, n=70, chunk=64),
    "lstm": dict(h=4, dt=0.1, n=70, chunk=64),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_step_host_control_flow_matches_whole_batch_step(cpu_doubles, name):
    c = CASES[name]
    h, dt, n = c["h"], c["dt"], c["n"]
    system = "quad" if name in ("quad", "autoregressive", "lstm") else name
    mode = name if name in ("autoregressive", "lstm") else "concurrent"
    w = dict(system=system, mode=mode, h=h, dt=dt)
    params = B.default_init(system, h, seed=4, mode=mode)
    spec = B.make_spec(w)
    case = B.make_case(w, n, 21, "cpu")
    a = T.FusedTrainStep(params, spec, n, lr=1e-4, device="cpu", distributed=False)
    b = T.FusedTrainStep(params, spec, n, lr=1e-4, device="cpu", distributed=False)
    # FusedTrainStep.step moves host tensors to the device; on this CPU double the "device" is the host
    a._dev = lambda x: x
    for it in range(2):
        la = a.step(case.get("in_state"), case["cur"], case.get("in_ref"), case.get("ref"), case.get("h0c0"))
        lb = b.step_host(case["cur"], ref=case.get("ref"), h0c0=case.get("h0c0"), target=case.get("target"),
                         chunk=c["chunk"])
        assert abs(float(la) - float(lb)) <= 2e-5 * abs(float(la)), (name, it, float(la), float(lb))
        assert rel_err(b.grad, a.grad) <= 2e-4, (name, it)
    assert rel_err(b.flat, a.flat) <= 1e-6
    sizes = sorted(set(_OracleRunner.made))
    assert n in sizes and (c["chunk"] in sizes or n <= c["chunk"])
    assert b.host_launches_per_step > 0
    # default chunk (patched to 128 here) and a single chunk go through the same code
    l0, g0 = b.value_and_grad_host(case["cur"], ref=case.get("ref"), h0c0=case.get("h0c0"), target=case.get("target"))
    l0, g0 = float(l0), g0.clone()
    l1, g1 = b.value_and_grad_host(case["cur"], ref=case.get("ref"), h0c0=case.get("h0c0"), target=case.get("target"),
                                   chunk=0)
    assert abs(l0 - float(l1)) <= 2e-5 * abs(l0) and rel_err(g1, g0) <= 2e-4


def test_bench_raw_e2e_arm_runs_its_check_and_reports(cpu_doubles):
    """bench.measure_e2e_raw end to end on the CPU doubles: cross-check, chunk selection, JSON fields"""
    _OracleRunner.lib = _capi.lib()
    n, h = 200, 10
    w = dict(B.WORKLOADS["quad_concurrent"], n=n)
    case = B.make_case(w, n, 1234, "cpu")
    params = B.default_init("quad", h)
    stepper = T.FusedTrainStep(params, B.make_spec(w), n, lr=1e-5, device="cpu", distributed=False)
    out = B.measure_e2e_raw(stepper, w, case, case, 3, 1, torch.device("cpu"), lambda: None)
    assert out["ok"] and out["value"] > 0 and out["h2d_bytes_per_step"] == n * (12 + 90) * 4
    assert out["check"]["grad_rel_l2"] <= 1e-4 and set(out["chunk_candidates_ms"]) >= {str(n), "64"}
    # a path that disagrees with the prepared-input path is reported, not timed
    bad = dict(case)
    bad["ref"] = case["ref"] + 1.0
    out = B.measure_e2e_raw(stepper, w, bad, case, 3, 1, torch.device("cpu"), lambda: None)
    assert out["ok"] is False and "loss_raw" in out["check"]


def test_step_host_async_equals_step_host(cpu_doubles):
    """the overlapped variant (loss read one step late, alternating staging sets) trains exactly like step_host"""
    n, h = 200, 10
    w = dict(B.WORKLOADS["quad_concurrent"], n=n)
    case = B.make_case(w, n, 7, "cpu")
    params = B.default_init("quad", h, seed=2)
    a = T.FusedTrainStep(params, B.make_spec(w), n, lr=1e-4, device="cpu", distributed=False)
    b = T.FusedTrainStep(params, B.make_spec(w), n, lr=1e-4, device="cpu", distributed=False)
    la = [float(a.step_host(case["cur"], ref=case["ref"], chunk=64)) for _ in range(4)]
    lb, prev = [], None
    for _ in range(4):
        hnd = b.step_host_async(case["cur"], ref=case["ref"], chunk=64)
        if prev is not None:
            lb.append(prev.item())
        prev = hnd
    lb.append(prev.item())
    assert la == lb and torch.equal(a.flat, b.flat)
    assert b._hs["sets"][0]["done"].recorded and b._hs["sets"][1]["done"].recorded


def test_captured_host_step_body_equals_step_host(cpu_doubles):
    """the body capture_host records (graph mode of value_and_grad_host: its own staging set, copy stream forked off and
    joined to the capturing stream, SGD update, loss into the pinned scalar) trains exactly like step_host"""
    n, h = 200, 10
    w = dict(B.WORKLOADS["quad_concurrent"], n=n)
    case = B.make_case(w, n, 9, "cpu")
    params = B.default_init("quad", h, seed=2)
    a = T.FusedTrainStep(params, B.make_spec(w), n, lr=1e-4, device="cpu", distributed=False)
    b = T.FusedTrainStep(params, B.make_spec(w), n, lr=1e-4, device="cpu", distributed=False)
    kw = dict(cur=case["cur"], ref=case["ref"], h0c0=None, target=None, norm=None, chunk=64)
    st = b._graph_state(n)
    la, lb = [], []
    for _ in range(3):
        la.append(float(a.step_host(case["cur"], ref=case["ref"], chunk=64)))
        b._host_step_body(kw, graph=True)
        lb.append(float(st["graph_loss_host"][0]))
    assert la == lb and torch.equal(a.flat, b.flat)
    assert st["sets"][0]["done"] is None and st["sets"][1]["done"] is None     # the eager staging sets stay untouched
