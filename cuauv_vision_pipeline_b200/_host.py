"""Shared plumbing of the reference-facing mirrors: accept what the reference functions accept
(numpy uint8 arrays, as handed to ModuleBase.process(), core/base.py:762-768) or CUDA tensors,
run on the default context, and hand back the same kind of object."""
import numpy as np
import torch

from .runtime import default_context


def is_device(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


def ctx_for(x):
    return default_context(x.device.index if is_device(x) else 0)


def to_device(ctx, x):
    if is_device(x):
        return x.contiguous()
    arr = np.asarray(x)
    if arr.dtype != np.uint8:
        raise TypeError("expected a uint8 image, got %s" % arr.dtype)
    return ctx.upload(arr)


def like_input(ctx, x, t):
    """Return `t` (device tensor) as numpy when the caller passed numpy."""
    return t if is_device(x) else ctx.download(t)
