"""numpy <-> CUDA tensor plumbing for the reference-shaped API.

numpy in -> numpy out (fp64, like the reference); torch CUDA tensor in -> torch CUDA tensor out
(fp64 or fp32, the tensor's dtype).  Compute always happens on the GPU: without CUDA these
helpers raise, they never compute on the host.
"""
from __future__ import annotations

import numpy as np
import torch


def is_tensor(x):
    return isinstance(x, torch.Tensor)


def any_tensor(*xs):
    return any(is_tensor(x) for x in xs)


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("model_predictive_control_b200 needs a CUDA device (B200); there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def pick_dtype(*xs):
    for x in xs:
        if is_tensor(x) and x.dtype in (torch.float32, torch.float64):
            return x.dtype
    return torch.float64


def to_dev(x, dtype=None, dev=None):
    """numpy / list / scalar / tensor -> CUDA tensor of ``dtype``."""
    dev = dev or device()
    if is_tensor(x):
        return x.to(device=dev, dtype=dtype or x.dtype)
    a = np.asarray(x)
    if a.dtype == object:
        raise ValueError("cannot convert object array")
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))
    return t.to(device=dev, dtype=dtype or torch.float64)


def back(t, as_numpy):
    if t is None:
        return None
    return t.detach().cpu().numpy() if as_numpy else t
