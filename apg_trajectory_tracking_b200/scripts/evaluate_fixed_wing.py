"""``FixedWingEvaluator`` of the reference (scripts/evaluate_fixed_wing.py:16-178) on the batched evaluation kernel:
``run_eval`` flies its ``nr_test`` random targets in ONE launch of ``apg_eval_fly_to_points``."""
import numpy as np
import torch

from .. import evaluate as EV, rollout as R, train as T


class FixedWingEvaluator:
    def __init__(self, controller, env, dt=0.01, horizon=1, render=0, thresh_div=10, thresh_stable=0.8, test_time=0,
                 **kwargs):
        self.controller, self.eval_env = controller, env
        self.dt, self.horizon, self.render = dt, horizon, render
        self.thresh_div, self.thresh_stable, self.test_time = thresh_div, thresh_stable, test_time
        self.des_speed = 11.5

    def fly_to_points(self, targets, max_steps=1000):
        """targets (N, K, 3): N simultaneous ``fly_to_point`` flights -> the dict of ``WingTargetEvaluator.fly``"""
        net, ds = self.controller.net, self.controller.dataset
        dev = next(net.parameters()).device
        spec = T.spec_for_net(net, "wing", self.horizon, self.eval_env.dt,
                              modified_params=getattr(self.eval_env.dynamics, "cfg", None))
        targets = torch.as_tensor(targets, dtype=torch.float32).to(dev)
        ev = EV.WingTargetEvaluator(spec, targets.shape[0], ds.mean, ds.std, getattr(ds, "dt", self.dt), dev)
        flat = R.flatten_params([p.detach() for p in net.parameters()]).to(dev).float().contiguous()
        out = ev.fly(flat, targets, steps=max_steps, thresh_div=self.thresh_div, thresh_stable=self.thresh_stable,
                     test_time=self.test_time)
        ctrl = self.controller
        take = getattr(ctrl, "take_every_x", 0)
        if take and ds is not None and hasattr(ctrl, "action_counter"):
            # self-play feed (network_wrapper.py:81-90): the kept calls' (state, target) go into the dataset's ring
            s, tg, _ = EV.wing_selfplay_samples(out, targets, take, ctrl.action_counter)
            s, tg = s.cpu().numpy(), tg.cpu().numpy()
            for i in range(len(s)):
                ds.get_and_add_eval_data(s[i].copy(), tg[i].copy(), add_to_dataset=True)
        if hasattr(ctrl, "action_counter"):
            ctrl.action_counter += int(out["n_steps"].sum())
        return out

    def run_eval(self, nr_test, return_dists=False, x_dist=50, x_std=5, printout=True):
        """evaluate_fixed_wing.py:133-178: targets [x_dist, U(-x_std, x_std), U(-x_std, x_std)] drawn in run order"""
        targets = np.zeros((nr_test, 1, 3))
        for i in range(nr_test):
            targets[i, 0] = [x_dist, *((np.random.rand(2) - .5) * 2 * x_std)]
        out = self.fly_to_points(targets)
        per_run = (out["div_target_sum"] / out["div_target_cnt"].clamp(min=1)).cpu().double().numpy()
        if printout:
            ns = out["n_steps"].cpu().double().numpy()
            print("Time not diverged: %3.2f (%3.2f)" % (ns.mean(), ns.std()))
            print("Average error (target): %3.2f (%3.2f)" % (per_run.mean(), per_run.std()))
        if return_dists:
            return per_run
        return EV.wing_eval_statistics(out["div_target_sum"], out["div_target_cnt"])
