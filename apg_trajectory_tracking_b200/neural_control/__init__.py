"""Mirror of the reference's ``neural_control`` import surface for the rollout hot path: same module
paths, class names, constructor signatures and parameter names, computed by the CUDA kernels of this package."""
