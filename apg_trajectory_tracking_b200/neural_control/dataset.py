"""Per-step featurizer of the recurrent train modes (reference: ``state_preprocessing`` in
``neural_control/dataset.py:207-220``): (N,12) quadrotor state -> (N,15) policy features
[vel, first two columns of the world->body matrix (row-major), body-frame velocity, body rates], differentiable,
one kernel forward / one backward.  The Dataset classes of the reference are host-side data preparation and stay
with the reference."""
from ..ops import quad_features


def state_preprocessing(drone_states):
    return quad_features(drone_states)
