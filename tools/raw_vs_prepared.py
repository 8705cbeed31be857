"""raw-sample mode vs prepared-input mode of the tcgen05 path at bench size: are actions / states / gradient the same
bits?  (diagnostic; the parity tests proper are tests/test_zz_new_paths_gpu.py tc3 / tc4)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from apg_trajectory_tracking_b200 import rollout as R, synthetic as SY, prepare as PR
import bench


def main():
    n, h, dt = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 10, 0.1
    dev = "cuda:0"
    case = SY.quad_case(n, h, dt, seed=1234, device=dev)
    flat = R.flatten_params(bench.default_init("quad", h)).to(dev)
    r = R.Rollout(R.RolloutSpec.quad_concurrent(h, dt), n, dev)
    prep = PR.prepare_quad(case["cur"].clone(), case["ref"].clone())
    for k in ("in_state", "cur", "in_ref", "ref"):
        print(k, "device prepare == host prepare:", bool(torch.equal(prep[k], case[k])),
              float((prep[k] - case[k]).abs().max()))
    out = {}
    for name, args in (("prepared", (prep["in_state"], prep["cur"], prep["in_ref"], prep["ref"])),
                       ("raw", (None, case["cur"], None, case["ref"])),
                       ("prepared2", (prep["in_state"], prep["cur"], prep["in_ref"], prep["ref"])),
                       ("raw2", (None, case["cur"], None, case["ref"]))):
        loss, st, ac = r.forward(flat, *args, want_states=True, want_actions=True)
        g = r.backward(1.0).clone()
        out[name] = (float(loss.item()), st.clone(), ac.clone(), g)
    torch.cuda.synchronize()
    base = out["prepared"]
    for name, (l, st, ac, g) in out.items():
        print(f"{name:10s} loss {l:.1f}  actions max abs diff {float((ac - base[2]).abs().max()):.3e}  states "
              f"{float((st - base[1]).abs().max()):.3e}  grad rel l2 {float((g - base[3]).norm() / base[3].norm()):.3e}")


main()
