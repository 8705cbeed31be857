"""``QuadEvaluator`` of the reference (scripts/evaluate_drone.py:28-298) on the batched evaluation kernel: the same
constructor and ``run_eval`` signature / return values, but the ``nr_test`` runs of one ``run_eval`` call are ONE
launch of ``apg_eval_rollout`` instead of ``nr_test`` x up to 251 policy calls on the host.  The controller's action
counter and the dataset's self-play slots end up as if the runs had been taken one after the other
(``evaluate.selfplay_samples``)."""
import os

import numpy as np
import torch

from .. import evaluate as EV, prepare as PR, rollout as R, train as T


class QuadEvaluator:
    def __init__(self, controller, environment, ref_length=5, max_drone_dist=0.1, render=0, dt=0.05, test_time=0,
                 speed_factor=.6, train_mode="concurrent", data_dir="data/traj_data_1", **kwargs):
        self.controller, self.eval_env = controller, environment
        self.horizon, self.max_drone_dist, self.render, self.dt = ref_length, max_drone_dist, render, dt
        self.action_counter = 0
        self.test_time, self.speed_factor, self.train_mode = test_time, speed_factor, train_mode
        self.data_dir = data_dir
        if hasattr(self.controller.net, "reset_hidden_state"):
            self.controller.net.reset_hidden_state()

    # ---- reference tables: Random.__init__ (trajectory/random_traj.py:29-36) for nr_test runs
    def load_tables(self, nr_test, device):
        """``nr_test`` random files of ``data_dir/{train,test}`` as ``load_prepare_trajectory`` picks and lays them
        out (generate_trajectory.py:566-603), +3 on z; files of one length each -> (nr_test, rows, 9)"""
        folder = os.path.join(self.data_dir, "test" if self.test_time else "train")
        names = sorted(os.listdir(folder))
        tabs = []
        for _ in range(nr_test):
            raw = np.load(os.path.join(folder, np.random.choice(names)))
            tabs.append(PR.reference_table(torch.as_tensor(raw, dtype=torch.float32).to(device), self.dt,
                                           self.speed_factor))
        if len({t.shape[0] for t in tabs}) != 1:
            raise ValueError("trajectory files of different lengths: pass tables= one length at a time")
        return torch.stack(tabs)

    def follow_tables(self, tables, max_nr_steps=200, thresh_stable=.4, thresh_div=3, init_states=None):
        """all rows of ``tables`` (T, rows, 9) as T simultaneous ``follow_trajectory("rand")`` runs -> the dict of
        ``TableEvaluator.follow`` (states, div, actions, n_steps)"""
        net = self.controller.net
        dev = next(net.parameters()).device
        dyn = self.eval_env.dynamics
        spec = T.spec_for_net(net, "quad", self.horizon, self.eval_env.dt, self.train_mode,
                              modified_params=getattr(dyn, "cfg", None))
        tables = tables.to(dev, torch.float32)
        ev = EV.TableEvaluator(spec, tables.shape[0], dev)
        flat = R.flatten_params([p.detach() for p in net.parameters()]).to(dev).float().contiguous()
        h0c0 = None
        if spec.net == "lstm":
            # the reference carries ONE hidden / cell state through its consecutive runs (rnn.py:45-48); here the T runs
            # are simultaneous: each starts from the net's current state, the net keeps the last run's final state
            hs, cs = net.hidden_state.detach().to(dev).float(), net.cell_state.detach().to(dev).float()
            h0c0 = torch.stack((hs[:1].expand(tables.shape[0], -1), cs[:1].expand(tables.shape[0], -1))).contiguous()
        out = ev.follow(flat, tables, init_states=init_states, steps=max_nr_steps, thresh_div=thresh_div,
                        thresh_stable=thresh_stable, test_time=self.test_time, h0c0=h0c0)
        if h0c0 is not None:
            net.hidden_state, net.cell_state = out["hc"][0, -1:].clone(), out["hc"][1, -1:].clone()
        self._feed_self_play(out, tables, thresh_div, thresh_stable)
        return out

    def _feed_self_play(self, out, tables, thresh_div, thresh_stable):
        ctrl = self.controller
        take = getattr(ctrl, "take_every_x", 0)
        total = int(out["n_steps"].sum())
        if take and getattr(ctrl, "dataset", None) is not None:
            s, r, _ = EV.selfplay_samples(out, tables, None, self.horizon, take, thresh_div, thresh_stable,
                                          self.test_time, ctrl.action_counter)
            s, r = s.cpu().numpy(), r.cpu().numpy()
            for i in range(len(s)):
                ctrl.dataset.get_and_add_eval_data(s[i].copy(), r[i].copy(), add_to_dataset=True)
        if hasattr(ctrl, "action_counter"):
            ctrl.action_counter += total

    def run_eval(self, reference="rand", nr_test=10, max_steps=251, thresh_div=1, thresh_stable=1, return_dict=False,
                 tables=None, **kwargs):
        """evaluate_drone.py:236-298; ``tables`` (nr_test, rows, 9) overrides the trajectory files"""
        if reference != "rand":
            raise NotImplementedError("only the table reference (\"rand\") of the trainers is on this path")
        np.random.seed(42)
        if nr_test == 0:
            return 0, 0
        dev = next(self.controller.net.parameters()).device
        if tables is None:
            tables = self.load_tables(nr_test, dev)
        out = self.follow_tables(tables, max_nr_steps=max_steps, thresh_stable=thresh_stable, thresh_div=thresh_div)
        stats = EV.eval_statistics(out["div"], out["n_steps"], thresh_div)
        suc_mean, suc_std, full_mean, full_std, div_mean, div_std = stats
        print("Average tracking error: %3.2f (%3.2f)" % (div_mean, div_std))
        if return_dict:
            ns = out["n_steps"].cpu().numpy()
            d = out["div"].cpu().numpy()
            stable = np.array([(d[i, :ns[i]] < thresh_div).sum() for i in range(len(ns))])
            return {"avg_tracking_error": full_mean, "std_tracking_error": full_std,
                    "ratio_stable": float((stable == ns[-1]).sum() / len(ns))}
        return stats
