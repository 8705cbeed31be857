"""Minimal end-to-end use of the drop-in surface: the quadrotor concurrent trainer of the reference
(scripts/train_drone.py, train_mode="concurrent") on synthetic polynomial references.

    python examples/train_quad_concurrent.py [n_samples] [epochs]

Raw (state, reference) samples -> QuadDataset (host-side prepare_data layouts) -> DataLoader -> TrainDrone.run_epoch,
whose mini-batch body is one fused forward + adjoint launch and torch.optim.SGD(momentum=0.9).step()."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from apg_trajectory_tracking_b200 import synthetic as SY                                   # noqa: E402
from apg_trajectory_tracking_b200.scripts.train_drone import TrainDrone                    # noqa: E402
from neural_control.dataset import QuadDataset                                             # noqa: E402
from neural_control.dynamics.quad_dynamics_flightmare import FlightmareDynamics            # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    epochs = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    horizon, dt = 10, 0.1
    raw = SY.quad_case(n, horizon, dt, seed=0)                 # raw drone states + absolute references
    data = QuadDataset(raw["cur"].numpy(), raw["ref"].numpy())
    config = dict(delta_t=dt, horizon=horizon, ref_dim=9, action_dim=4, state_size=12, batch_size=512, system="quad",
                  learning_rate_controller=1e-5, train_mode="concurrent")
    trainer = TrainDrone(FlightmareDynamics(), FlightmareDynamics(), config)
    torch.manual_seed(0)
    trainer.initialize_model(state_data=data)
    for epoch in range(epochs):
        print(f"epoch {epoch}: mean batch loss {trainer.run_epoch(epoch=epoch):.3f}")


if __name__ == "__main__":
    main()
