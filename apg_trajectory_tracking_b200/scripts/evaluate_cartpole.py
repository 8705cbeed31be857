"""``Evaluator.evaluate_in_environment`` of the reference (scripts/evaluate_cartpole.py:40-262, state-based
controller) on the batched evaluation kernel: the ``nr_iters`` runs are ONE launch of ``apg_eval_cartpole``."""
import numpy as np
import torch

from .. import evaluate as EV, rollout as R, train as T


class Evaluator:
    def __init__(self, controller, eval_env, eval_dyn=None, **kwargs):
        self.controller, self.eval_env, self.eval_dyn = controller, eval_env, eval_dyn
        self.initialize_straight = 1

    def evaluate_in_environment(self, nr_iters=1, max_steps=250, render=False, burn_in_steps=50, return_success=0,
                                init_states=None):
        """evaluate_cartpole.py:78-262; ``init_states`` (nr_iters, 4) overrides the reference's start (all zeros when
        ``initialize_straight``, else ``CartPoleEnv._reset_upright`` draws)"""
        if nr_iters == 0:
            return 0, 0, []
        net = self.controller.net
        dev = next(net.parameters()).device
        if init_states is None:
            if self.initialize_straight:
                init_states = np.zeros((nr_iters, 4))
            else:
                init_states = np.stack([np.array(self.eval_env._reset_upright()) for _ in range(nr_iters)])
        spec = T.spec_for_net(net, "cartpole", self.controller.horizon, self.eval_env.dt,
                              modified_params=getattr(self.eval_env.dynamics, "cfg", None))
        ev = EV.CartpoleBalanceEvaluator(spec, nr_iters, dev)
        flat = R.flatten_params([p.detach() for p in net.parameters()]).to(dev).float().contiguous()
        out = ev.balance(flat, torch.as_tensor(init_states, dtype=torch.float32).to(dev), steps=max_steps,
                         thresh_div=self.eval_env.thresh_div, burn_in_steps=burn_in_steps, want=("states",))
        ns = out["n_steps"].cpu().numpy()
        states = out["states"].cpu().numpy()
        velocities = np.concatenate([np.abs(states[i, :ns[i], 1]) for i in range(nr_iters)])
        success = (ns - 1).astype(np.float64)
        res = {"mean_vel": float(np.mean(velocities)), "std_vel": float(np.std(velocities)),
               "mean_stable": float(np.mean(success)), "std_stable": float(np.std(success))}
        print("Average velocity: %3.2f (%3.2f)" % (res["mean_vel"], res["std_vel"]))
        print("Average success: %3.2f (%3.2f)" % (res["mean_stable"], res["std_stable"]))
        if return_success:
            return success, velocities.tolist()
        return res
