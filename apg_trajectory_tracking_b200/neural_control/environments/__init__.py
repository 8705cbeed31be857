import torch


def compute_device():
    """the device the single-drone environments run their dynamics step on (CUDA only: there is no CPU path)"""
    if not torch.cuda.is_available():
        raise RuntimeError("no CUDA device: the environments step through the CUDA dynamics op")
    return torch.device("cuda", torch.cuda.current_device())
